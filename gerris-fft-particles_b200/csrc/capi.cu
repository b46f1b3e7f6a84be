/* capi.cu -- device context and the C-ABI entry points of include/gfsb200.h.
 *
 * HBM layout per context (one context per GPU / rank):
 *   tree      child0 (int32, -1 leaf / -2 destroyed), neighbor [n][2*dim] int32,
 *             level + info (uint8), leaf_vtx [n][2^dim] int32, vertex CSR
 *             (vtx_off int32, vtx_cell int32, vtx_w fp64)      -- re-uploaded per adapt
 *   field     U,V,W (+alpha, mu) fp64 SoA in flat-tree order    -- mirrored per step
 *   derived   vtx_val [n_vertices][4] fp64, vort [n_cells][4] fp64 -- rebuilt per field update
 *   particles x,y,z,vx,vy,vz,mass,volume fp64 SoA (+ alternate buffers for the
 *             sort/cull permutation), id uint32, optional fx,fy,fz, cell
 *   deposit   [1 + dim][n_cells] fp64 (void fraction, force components)
 *
 * All work is issued on the context's own non-blocking stream.  There is no
 * CPU fallback anywhere in this file: without a usable sm_100 device
 * gfsb200_ctx_create fails with GFSB200_ERR_CUDA.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include "gfsb200_internal.h"
#include "device_types.cuh"

extern "C" {
void gfsb200_launch_cell_pass (const DevTree *, const DevField *, int, cudaStream_t, cudaStream_t,
			       cudaEvent_t, cudaEvent_t);
void gfsb200_launch_vertex_values (const DevTree *, const DevField *, int, cudaStream_t);
void gfsb200_launch_convective (const DevTree *, const DevField *, int, cudaStream_t);
int gfsb200_launch_step (const DevTree *, const DevField *, const DevParticles *, const DevStep *,
			 int, int, int, int, cudaStream_t, const DevDeposit *);   /* (.., record, min blocks/SM, mode, SMs, stream, fused deposit) */
void gfsb200_launch_advect (const DevTree *, const DevField *, const DevParticles *, double, int,
			    cudaStream_t);
void gfsb200_launch_locate (const DevTree *, int64_t, const double *, const double *,
			    const double *, int32_t *, cudaStream_t);
void gfsb200_launch_interpolate (const DevTree *, const DevField *, int64_t, const double *,
				 const double *, const double *, double *, double *, double *,
				 cudaStream_t);
void gfsb200_launch_corner_values (const DevTree *, const DevField *, int, int64_t, const int32_t *,
				   double *, cudaStream_t);
void gfsb200_launch_deposit (const DevTree *, const DevField *, const DevParticles *, const DevStep *,
			     int, const DevDeposit *, int, const int32_t *, const uint8_t *, cudaStream_t);
void gfsb200_launch_deposit_smoothed (const DevTree *, const DevField *, const DevParticles *,
				      const DevStep *, double, const gfsb200_kernel *, double *, double *,
				      double *, double *, cudaStream_t);
void gfsb200_launch_gather3 (int64_t, const int32_t *, const double *, const double *, const double *,
			     double *, double *, double *, cudaStream_t);
void gfsb200_launch_gather (int64_t, const int32_t *, int, const double * const *, double * const *,
			    const uint32_t *, uint32_t *, cudaStream_t);
void gfsb200_launch_particle_bc (const DevTree *, const DevParticles *, int, const int32_t *,
				 const double *, uint8_t *, int *, cudaStream_t);
void gfsb200_launch_iota (int64_t, int32_t *, cudaStream_t);
void gfsb200_launch_iota_u32 (int64_t, uint32_t *, uint32_t, cudaStream_t);
void gfsb200_launch_sort_keys (int64_t, const int32_t *, uint32_t *, uint32_t, cudaStream_t);
void gfsb200_launch_owner_keys (int64_t, const uint32_t *, uint32_t, const uint8_t *, uint32_t *, uint32_t, cudaStream_t);
void gfsb200_launch_inside_flags (int64_t, const int32_t *, uint8_t *, cudaStream_t);
void gfsb200_launch_outside_clear (int64_t, const int32_t *, uint8_t *, int *, cudaStream_t);
cudaError_t gfsb200_cub_sort_pairs (void *, size_t *, const uint32_t *, uint32_t *, const int32_t *,
				    int32_t *, int64_t, int, cudaStream_t);
cudaError_t gfsb200_cub_select_flagged (void *, size_t *, const int32_t *, const uint8_t *,
					int32_t *, int32_t *, int64_t, cudaStream_t);
}

#include "ctx_internal.cuh"

/* number of this library's own kernels launched so far in the process (CUB's sort / select
   kernels are not counted) */
extern "C" {
long long gfsb200_launch_counter = 0;
int64_t gfsb200_kernel_launches (void) { return gfsb200_launch_counter; }
}

template <typename Tp>
static int dev_alloc_copy (Tp ** dst, const Tp * src, size_t n, cudaStream_t st)
{
  *dst = NULL;
  CK (cudaMalloc ((void **) dst, (n ? n : 1)*sizeof (Tp)));
  if (n && src)
    CK (cudaMemcpyAsync (*dst, src, n*sizeof (Tp), cudaMemcpyHostToDevice, st));
  return GFSB200_OK;
}

/* Device scratch of the batched queries: ONE pooled allocation per context that only grows, carved
 * up by a bump pointer -- an output event that fires every step pays no cudaMalloc / cudaFree (each
 * of which synchronises the device). */
struct Scratch {
  gfsb200_ctx * c;
  size_t used;
  Scratch (gfsb200_ctx * ctx) : c (ctx), used (0) {}
  static size_t pad (size_t b) { return (b + 255) & ~(size_t) 255; }
  /* total: the sum of the padded sizes about to be taken */
  int reserve (size_t total)
  {
    if (total <= c->scratch_bytes) return GFSB200_OK;
    CK (cudaStreamSynchronize (c->stream));
    cudaFree (c->scratch);
    c->scratch = NULL; c->scratch_bytes = 0;
    total += total/4;
    CK (cudaMalloc ((void **) &c->scratch, total));
    c->scratch_bytes = total;
    return GFSB200_OK;
  }
  template <typename Tp> Tp * take (size_t n)
  {
    Tp * p = reinterpret_cast<Tp *> (reinterpret_cast<char *> (c->scratch) + used);
    used += pad (n*sizeof (Tp));
    return p;
  }
};

static void free_tree (gfsb200_ctx * c)
{
  cudaFree (c->d_child0); cudaFree (c->d_neighbor); cudaFree (c->d_la_slot);
  cudaFree (c->d_vtx_off); cudaFree (c->d_vtx_cell); cudaFree (c->d_leaf_vtx); cudaFree (c->d_parent);
  c->d_parent = NULL;
  cudaFree (c->d_solid_a); c->d_solid_a = NULL;
  cudaFree (c->d_solid_s); c->d_solid_s = NULL;
  cudaFree (c->d_level); cudaFree (c->d_info); cudaFree (c->d_vtx_w); cudaFree (c->d_vtx_wuni);
  c->d_child0 = c->d_neighbor = c->d_la_slot = c->d_vtx_off = c->d_vtx_cell = c->d_leaf_vtx = NULL;
  c->d_level = c->d_info = NULL; c->d_vtx_w = c->d_vtx_wuni = NULL;
  cudaFree (c->F.vtx_val); cudaFree (c->F.vort); cudaFree (c->F.nodata_flag);
  c->F.vtx_val = c->F.vort = NULL; c->F.nodata_flag = NULL;
  for (int i = 0; i < 5; i++) { cudaFree (c->d_field[i]); c->d_field[i] = NULL; }
  for (int i = 0; i < 3; i++) { cudaFree (c->d_prev[i]); c->d_prev[i] = NULL; }
  cudaFree (c->F.vtx_prev); cudaFree (c->F.acc);
  c->F.vtx_prev = c->F.acc = NULL;
  c->have_prev = c->acc_valid = false;
  cudaFree (c->deposit_buf[0]); cudaFree (c->deposit_buf[1]);
  cudaFree (c->knorm); c->knorm = NULL; c->knorm_cap = c->knorm_n = 0;
  c->deposit = c->deposit_buf[0] = c->deposit_buf[1] = NULL; c->deposit_count = 0;
  c->dep_which = c->dep_result = 0;
  c->have_tree = c->have_field = false;
}

static void free_particles (gfsb200_ctx * c)
{
  for (int b = 0; b < 2; b++) {
    for (int k = 0; k < NCOL; k++) { cudaFree (c->col[b][k]); c->col[b][k] = NULL; }
    cudaFree (c->id[b]); c->id[b] = NULL;
  }
  for (int k = 0; k < 3; k++) { cudaFree (c->force[k]); c->force[k] = NULL; }
  cudaFree (c->cell); cudaFree (c->perm); cudaFree (c->perm2); cudaFree (c->key); cudaFree (c->key2);
  cudaFree (c->flag); cudaFree (c->cub_tmp);
  c->cell = c->perm = c->perm2 = NULL; c->key = c->key2 = NULL; c->flag = NULL; c->cub_tmp = NULL;
  c->cub_tmp_bytes = 0;
  c->n = c->cap = c->aux_cap = 0;
  cudaFree (c->esc_count); cudaFree (c->esc_idx); cudaFree (c->esc_old);
  c->esc_count = NULL; c->esc_idx = NULL; c->esc_old = NULL; c->esc_cap = 0; c->esc_armed = false;
}

extern "C" int gfsb200_device_count (void)
{
  int n = 0;
  if (cudaGetDeviceCount (&n) != cudaSuccess) {
    cudaGetLastError ();
    return 0;
  }
  return n;
}

extern "C" int gfsb200_ctx_create (int device, gfsb200_ctx ** out)
{
  if (!out)
    return gfsb200_fail (GFSB200_ERR_ARG, "ctx_create: null output");
  *out = NULL;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount (&ndev);
  if (e != cudaSuccess || ndev == 0)
    return gfsb200_fail (GFSB200_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path",
			 cudaGetErrorString (e));
  if (device < 0 || device >= ndev)
    return gfsb200_fail (GFSB200_ERR_ARG, "ctx_create: device %d out of range (%d devices)", device, ndev);
  CK (cudaSetDevice (device));
  cudaDeviceProp prop;
  CK (cudaGetDeviceProperties (&prop, device));
  if (prop.major != 10)
    return gfsb200_fail (GFSB200_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
			 device, prop.major, prop.minor);
  gfsb200_ctx * c = new gfsb200_ctx ();
  memset (&c->T, 0, sizeof c->T);
  memset (&c->F, 0, sizeof c->F);
  c->device = device;
  c->n_sm = prop.multiProcessorCount;
  c->have_tree = c->have_field = c->own_field = false;
  c->d_child0 = c->d_neighbor = c->d_la_slot = c->d_vtx_off = c->d_vtx_cell = c->d_leaf_vtx = NULL;
  c->d_level = c->d_info = NULL; c->d_vtx_w = c->d_vtx_wuni = NULL; c->d_parent = NULL;
  c->d_solid_a = c->d_solid_s = NULL;
  c->esc_count = NULL; c->esc_idx = NULL; c->esc_old = NULL; c->esc_cap = 0; c->esc_armed = false;
  c->mark_outside = false;
  c->forces_recorded = false;
  for (int i = 0; i < 5; i++) c->d_field[i] = NULL;
  for (int i = 0; i < 3; i++) c->d_prev[i] = NULL;
  c->have_prev = c->acc_valid = false;
  c->n = c->cap = c->aux_cap = 0; c->cur = 0;
  for (int b = 0; b < 2; b++) { for (int k = 0; k < NCOL; k++) c->col[b][k] = NULL; c->id[b] = NULL; }
  for (int k = 0; k < 3; k++) c->force[k] = NULL;
  c->cell = c->perm = c->perm2 = NULL; c->key = c->key2 = NULL; c->flag = NULL;
  c->d_count = NULL; c->cub_tmp = NULL; c->cub_tmp_bytes = 0; c->d_ptr_table = NULL;
  c->deposit = c->deposit_buf[0] = c->deposit_buf[1] = NULL; c->deposit_count = 0;
  c->dep_which = c->dep_result = 0;
  c->comm = NULL; c->tree_generation = 0;
  c->scratch = NULL; c->scratch_bytes = 0;
  memset (c->hp_col, 0, sizeof c->hp_col);
  c->hp_chunk = 0; c->hp_h2d = c->hp_d2h = NULL;
  c->ev_used = 0; c->timing = true; c->timer_every = 1; c->timer_calls = 0;
  c->step_minb = getenv ("GFSB200_STEP_MINB") ? atoi (getenv ("GFSB200_STEP_MINB")) : 3;
  c->step_mode = getenv ("GFSB200_STEP_MODE") ? atoi (getenv ("GFSB200_STEP_MODE")) : -1;
  if (cudaStreamCreateWithFlags (&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags (&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags (&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags (&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaMalloc ((void **) &c->d_count, sizeof (int32_t)) != cudaSuccess ||
      cudaMalloc ((void **) &c->d_ptr_table, 2*NCOL*sizeof (double *)) != cudaSuccess) {
    delete c;
    return gfsb200_fail (GFSB200_ERR_CUDA, "ctx_create: %s", cudaGetErrorString (cudaGetLastError ()));
  }
  *out = c;
  return GFSB200_OK;
}

extern "C" void gfsb200_ctx_destroy (gfsb200_ctx * c)
{
  if (!c) return;
  cudaSetDevice (c->device);
  cudaStreamSynchronize (c->stream);
  if (c->comm) gfsb200_comm_detach (c->comm);
  free_tree (c);
  free_particles (c);
  cudaFree (c->scratch);
  for (int s = 0; s < 3; s++)
    for (int k = 0; k < NCOL; k++) cudaFree (c->hp_col[s][k]);
  if (c->hp_h2d) {
    cudaStreamDestroy (c->hp_h2d); cudaStreamDestroy (c->hp_d2h);
    for (int s = 0; s < 3; s++) {
      cudaEventDestroy (c->hp_in[s]); cudaEventDestroy (c->hp_done[s]); cudaEventDestroy (c->hp_out[s]);
    }
  }
  cudaFree (c->d_count); cudaFree (c->d_ptr_table);
  for (size_t i = 0; i < c->ev.size (); i++) cudaEventDestroy (c->ev[i]);
  cudaEventDestroy (c->ev_fork); cudaEventDestroy (c->ev_join);
  cudaStreamDestroy (c->aux_stream);
  cudaStreamDestroy (c->stream);
  delete c;
}

extern "C" void * gfsb200_ctx_stream (gfsb200_ctx * c) { return c ? (void *) c->stream : NULL; }

extern "C" int gfsb200_ctx_synchronize (gfsb200_ctx * c)
{
  if (!c) return gfsb200_fail (GFSB200_ERR_ARG, "null context");
  CK (cudaSetDevice (c->device));
  CK (cudaStreamSynchronize (c->stream));
  return GFSB200_OK;
}

/* ------------------------------------------------------------------ */

extern "C" int gfsb200_upload_tree (gfsb200_ctx * c, const gfsb200_tree * t)
{
  if (!c || !t) return gfsb200_fail (GFSB200_ERR_ARG, "upload_tree: null argument");
  if (!t->finalized || !t->leaf_vtx)
    return gfsb200_fail (GFSB200_ERR_STATE, "upload_tree: tree needs finalize + build_stencils");
  if (t->n_roots > GFSB200_MAX_DEV_ROOTS)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "upload_tree: more than %d root cells", GFSB200_MAX_DEV_ROOTS);
  CK (cudaSetDevice (c->device));
  CK (cudaStreamSynchronize (c->stream));
  if (c->comm) gfsb200_comm_tree_changed (c->comm);   /* waits for its stream; ownership and peer pointers lapse */
  free_tree (c);
  c->tree_generation++;

  const int32_t n = t->n_cells;
  const int nc = t->nchild;
  std::vector<int32_t> child0 (n);
  std::vector<uint8_t> info (n);
  bool any_destroyed = false;
  int32_t leaf_lo = -1, leaf_hi = 0;
  for (int32_t i = 0; i < n; i++) {
    child0[i] = (t->flags[i] & GFSB200_CELL_DESTROYED) ? CHILD_DESTROYED :
      (t->child0[i] < 0 ? CHILD_LEAF : t->child0[i]);
    int k = t->parent[i] < 0 ? 0 : i - t->child0[t->parent[i]];
    unsigned regular = (t->flags[i] & GFSB200_CELL_LEAF) ? CELL_REGULAR : 0;
    if (t->solid_a && t->solid_a[i] != 1.)
      regular = 0;                    /* a mixed cell may have closed faces (gfs_cell_face) */
    else if (t->solid_cm && t->solid_cm[3*i] == t->solid_cm[3*i])
      regular = 0;
    for (int d = 0; d < t->ndir && regular; d++) {
      int32_t nb = t->neighbor[(int64_t) i*t->ndir + d];
      if (nb < 0 || t->level[nb] != t->level[i] || !(t->flags[nb] & GFSB200_CELL_LEAF))
	regular = 0;
    }
    info[i] = (uint8_t) ((t->flags[i] & 7) | regular | (k << 4));
    if ((t->flags[i] & GFSB200_CELL_LEAF) && !(t->flags[i] & GFSB200_CELL_BOUNDARY)) {
      if (leaf_lo < 0) leaf_lo = i;
      leaf_hi = i + 1;
    }
    if ((t->flags[i] & GFSB200_CELL_DESTROYED) && !(t->flags[i] & GFSB200_CELL_BOUNDARY))
      any_destroyed = true;
  }
  std::vector<double> wuni (t->n_vertices ? t->n_vertices : 1);
  for (int32_t v = 0; v < t->n_vertices; v++) {
    const int32_t b = t->vtx_off[v], e = t->vtx_off[v + 1];
    double w = e > b ? t->vtx_w[b] : 0.;
    for (int32_t j = b + 1; j < e; j++)
      if (t->vtx_w[j] != w) { w = NAN; break; }
    wuni[v] = w;
  }
  int r;
  if ((r = dev_alloc_copy (&c->d_child0, child0.data (), n, c->stream))) return r;
  if ((r = dev_alloc_copy (&c->d_info, info.data (), n, c->stream))) return r;
  if ((r = dev_alloc_copy (&c->d_level, (const uint8_t *) t->level, n, c->stream))) return r;
  if ((r = dev_alloc_copy (&c->d_neighbor, (const int32_t *) t->neighbor, (size_t) n*t->ndir, c->stream))) return r;
  if ((r = dev_alloc_copy (&c->d_la_slot, (const int32_t *) t->la_slot, t->la_size, c->stream))) return r;
  if ((r = dev_alloc_copy (&c->d_vtx_off, (const int32_t *) t->vtx_off, (size_t) t->n_vertices + 1, c->stream))) return r;
  const size_t ne = t->vtx_off[t->n_vertices];
  if ((r = dev_alloc_copy (&c->d_vtx_cell, (const int32_t *) t->vtx_cell, ne, c->stream))) return r;
  if ((r = dev_alloc_copy (&c->d_vtx_w, (const double *) t->vtx_w, ne, c->stream))) return r;
  if ((r = dev_alloc_copy (&c->d_vtx_wuni, (const double *) wuni.data (), (size_t) t->n_vertices, c->stream))) return r;
  if ((r = dev_alloc_copy (&c->d_leaf_vtx, (const int32_t *) t->leaf_vtx, (size_t) n*nc, c->stream))) return r;
  if ((r = dev_alloc_copy (&c->d_parent, (const int32_t *) t->parent, (size_t) n, c->stream))) return r;
  if (t->solid_a && (r = dev_alloc_copy (&c->d_solid_a, (const double *) t->solid_a, (size_t) n, c->stream))) return r;
  if (t->solid_s && (r = dev_alloc_copy (&c->d_solid_s, (const double *) t->solid_s, (size_t) n*t->ndir, c->stream))) return r;
  CK (cudaStreamSynchronize (c->stream));   /* host staging vectors go out of scope */

  DevTree & T = c->T;
  memset (&T, 0, sizeof T);
  T.dim = t->dim; T.n_cells = n; T.n_roots = t->n_roots; T.n_box_roots = t->n_box_roots;
  T.root_level = t->root_level;
  T.top_levels = t->complete_level - t->root_level;
  /* Morton keys of the arithmetic top are built from 10 (3D) / 16 (2D) bits per axis */
  if (T.top_levels > (t->dim == 3 ? 10 : 15)) T.top_levels = t->dim == 3 ? 10 : 15;
  T.top_start = t->level_start[T.top_levels];
  T.root_size = ldexp (1., -t->root_level);
  T.top_h = ldexp (1., -(t->root_level + T.top_levels));
  T.top_inv_h = ldexp (1., t->root_level + T.top_levels);
  T.la_inv_h = 1./t->la_h;
  for (int rr = 0; rr < t->n_roots; rr++)
    for (int a = 0; a < 3; a++)
      T.root_pos[rr][a] = t->pos[3*rr + a];
  for (int a = 0; a < 3; a++) { T.la_min[a] = t->la_min[a]; T.la_n[a] = t->la_n[a]; }
  T.la_h = t->la_h;
  T.single_box = t->la_size == 1 && t->la_slot[0] == 0;
  T.has_destroyed = any_destroyed;
  c->leaf_lo = leaf_lo < 0 ? 0 : leaf_lo;
  c->leaf_hi = leaf_hi;
  T.la_slot = c->d_la_slot;
  T.child0 = c->d_child0; T.neighbor = c->d_neighbor; T.level = c->d_level; T.info = c->d_info;
  T.n_vertices = t->n_vertices;
  T.vtx_off = c->d_vtx_off; T.vtx_cell = c->d_vtx_cell; T.vtx_w = c->d_vtx_w; T.leaf_vtx = c->d_leaf_vtx;
  T.vtx_wuni = c->d_vtx_wuni;
  T.parent = c->d_parent;
  T.solid_a = c->d_solid_a;            /* NULL when the tree has no mixed cell */
  T.solid_s = c->d_solid_s;
  for (int rr = 0; rr < GFSB200_MAX_DEV_ROOTS; rr++)
    for (int d = 0; d < 6; d++)
      T.periodic[rr][d] = rr < t->n_roots ? (signed char) t->periodic[rr][d] : -1;
  T.lattice_n1 = 0;
  /* (a one-slot locate array means one GfsBox and no GfsBoundary root: hull cells have NULL
     neighbours, so "all neighbours are same-level leaves" == interior on such a tree) */
  /* The lattice flavour of locate () lets the locate-array slot test stand for the root test of
     ftt_cell_locate as well: that needs the slot to BE the box -- same lower corner, same size --
     with corners that are exact in fp64 (they are for dyadic boxes; anything else keeps the
     general path). */
  bool slot_is_box = T.la_h == T.root_size;
  for (int a = 0; a < t->dim; a++) {
    const double c = T.root_pos[0][a], half = 0.5*T.root_size, lo = c - half, hi = c + half;
    slot_is_box = slot_is_box && T.la_min[a] == lo && lo + half == c && hi - half == c && lo + T.root_size == hi;
  }
  T.slot_is_box = T.single_box && t->n_roots == 1 && slot_is_box;
  if (t->lattice_level >= 0 && t->lattice_level - t->root_level == T.top_levels && T.single_box &&
      t->n_roots == 1 && !any_destroyed && slot_is_box && !getenv ("GFSB200_NO_LATTICE"))
    T.lattice_n1 = (1 << T.top_levels) + 1;
  /* Interior vertices of a lattice tree: check once, on the host, that every one of them
     carries the same stencil shape (the 2^dim leaves around it, equal weights, one common
     order), so that the cell pass can address them arithmetically instead of through the
     CSR tables.  Any deviation keeps the tables. */
  T.lattice_pattern = -1;
  if (T.lattice_n1 > 0 && T.top_levels >= 2 && !t->solid_a && !getenv ("GFSB200_NO_LATTICE_PATTERN")) {
    const int dim = t->dim, nc = 1 << dim, n1 = T.lattice_n1, nn = n1 - 1;
    double wexp = NAN;                /* the common weight (2^-dim up to the rounding of the
					 reference's sequential normalisation) */
    int pattern = -1;
    bool ok = true;
    for (int k = dim == 3 ? 1 : 0; ok && k < (dim == 3 ? nn : 1); k++)
      for (int j = 1; ok && j < nn; j++)
	for (int i = 1; ok && i < nn; i++) {
	  const int v = (k*n1 + j)*n1 + i;
	  const int b = t->vtx_off[v], e = t->vtx_off[v + 1];
	  if (pattern < 0) wexp = wuni[v];
	  if (e - b != nc || !(wuni[v] == wexp)) { ok = false; break; }
	  int pat = 0;
	  for (int q = 0; q < nc; q++) {
	    const int cell = t->vtx_cell[b + q];
	    if (cell < T.top_start || t->child0[cell] >= 0) { ok = false; break; }
	    const unsigned key = (unsigned) (cell - T.top_start);
	    int kx, ky, kz = 0;
	    if (dim == 3) {
	      kx = gfsb200_compact3 (key); ky = ~gfsb200_compact3 (key >> 1) & (nn - 1);
	      kz = ~gfsb200_compact3 (key >> 2) & (nn - 1);
	    }
	    else {
	      kx = gfsb200_compact2 (key); ky = ~gfsb200_compact2 (key >> 1) & (nn - 1);
	    }
	    const int bx = kx - (i - 1), by = ky - (j - 1), bz = dim == 3 ? kz - (k - 1) : 0;
	    if ((bx | by | bz) & ~1) { ok = false; break; }
	    pat |= (bx | by << 1 | bz << 2) << (dim*q);
	  }
	  if (pattern < 0) pattern = pat;
	  else if (pat != pattern) ok = false;
	}
    if (ok && pattern >= 0) {
      T.lattice_pattern = pattern;
      T.lattice_w = wexp;
    }
    /* 3D: can lattice_cell_pass_kernel tile the tree with 8^3 bricks?  It assumes that the
       complete level holds exactly the nn^3 box leaves and that a leaf has a NULL neighbour
       exactly where it touches the hull. */
    if (T.lattice_pattern >= 0 && dim == 3 && nn >= 16 && nn % 8 == 0 &&
	(int64_t) T.top_start + (int64_t) nn*nn*nn == n) {
      for (int32_t cell = T.top_start; cell < n && ok; cell++) {
	const unsigned key = (unsigned) (cell - T.top_start);
	const int k3[3] = { (int) gfsb200_compact3 (key), (int) (~gfsb200_compact3 (key >> 1) & (nn - 1)),
			    (int) (~gfsb200_compact3 (key >> 2) & (nn - 1)) };
	if ((info[cell] & (GFSB200_CELL_LEAF | GFSB200_CELL_BOUNDARY)) != GFSB200_CELL_LEAF) ok = false;
	for (int d = 0; d < 6 && ok; d++) {
	  const int32_t nbr = t->neighbor[(int64_t) cell*6 + d];
	  const bool at_hull = (d & 1) ? k3[d >> 1] == 0 : k3[d >> 1] == nn - 1;
	  if (at_hull ? nbr >= 0 : (nbr < T.top_start || t->child0[nbr] >= 0)) ok = false;
	}
      }
      T.lattice_bricks = ok;
    }
    if (getenv ("GFSB200_DEBUG"))
      fprintf (stderr, "gfsb200: lattice n1 = %d, interior vertex pattern %#x (w = %.17g), bricks %d\n",
	       T.lattice_n1, T.lattice_pattern, wexp, T.lattice_bricks);
  }

  memset (&c->F, 0, sizeof c->F);
  const int vs = t->dim == 3 ? 4 : 2, ws = t->dim == 3 ? 4 : 1;
  CK (cudaMalloc ((void **) &c->F.vtx_val, (size_t) (t->n_vertices ? t->n_vertices : 1)*vs*sizeof (double)));
  CK (cudaMalloc ((void **) &c->F.vort, (size_t) n*ws*sizeof (double)));
  CK (cudaMemsetAsync (c->F.vort, 0, (size_t) n*ws*sizeof (double), c->stream));
  CK (cudaMalloc ((void **) &c->F.nodata_flag, sizeof (int)));
  CK (cudaMemsetAsync (c->F.nodata_flag, 0xff, sizeof (int), c->stream));  /* -1: no epoch (they start at 1) */
  c->deposit_count = (int64_t) (1 + t->dim)*n;
  CK (cudaMalloc ((void **) &c->deposit_buf[0], (size_t) c->deposit_count*sizeof (double)));
  CK (cudaMemsetAsync (c->deposit_buf[0], 0, (size_t) c->deposit_count*sizeof (double), c->stream));
  c->deposit = c->deposit_buf[0];
  c->have_tree = true;
  c->have_field = false;
  return GFSB200_OK;
}

extern "C" int gfsb200_refresh_field (gfsb200_ctx * c)
{
  if (!c || !c->have_tree || !c->F.u[0])
    return gfsb200_fail (GFSB200_ERR_STATE, "refresh_field: no tree/field resident");
  CK (cudaSetDevice (c->device));
  /* a new epoch instead of clearing the flag (DevField.nodata_epoch) */
  c->F.nodata_epoch = c->F.nodata_epoch >= 0x7ffffff0 ? 1 : c->F.nodata_epoch + 1;
  gfsb200_launch_cell_pass (&c->T, &c->F, c->n_sm, c->stream, c->aux_stream, c->ev_fork, c->ev_join);
  CK (cudaGetLastError ());
  c->have_field = true;
  c->acc_valid = false;
  return GFSB200_OK;
}

extern "C" int gfsb200_upload_field_prev (gfsb200_ctx * c, const double * un, const double * vn,
					  const double * wn)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "upload_field_prev: upload a tree first");
  if (!un || !vn || (c->T.dim == 3 && !wn))
    return gfsb200_fail (GFSB200_ERR_ARG, "upload_field_prev: missing velocity component");
  CK (cudaSetDevice (c->device));
  const double * src[3] = { un, vn, wn };
  const size_t bytes = (size_t) c->T.n_cells*sizeof (double);
  for (int i = 0; i < c->T.dim; i++) {
    if (!c->d_prev[i]) CK (cudaMalloc ((void **) &c->d_prev[i], bytes));
    CK (cudaMemcpyAsync (c->d_prev[i], src[i], bytes, cudaMemcpyHostToDevice, c->stream));
  }
  const int vs = c->T.dim == 3 ? 4 : 2;
  if (!c->F.vtx_prev)
    CK (cudaMalloc ((void **) &c->F.vtx_prev, (size_t) (c->T.n_vertices ? c->T.n_vertices : 1)*vs*sizeof (double)));
  if (!c->F.nodata_flag)
    return gfsb200_fail (GFSB200_ERR_STATE, "upload_field_prev: upload the current field first");
  for (int i = 0; i < 3; i++) c->F.uprev[i] = c->d_prev[i];
  DevField tmp = c->F;
  for (int i = 0; i < 3; i++) tmp.u[i] = c->d_prev[i];
  tmp.vtx_val = c->F.vtx_prev;
  gfsb200_launch_vertex_values (&c->T, &tmp, c->n_sm, c->stream);
  CK (cudaGetLastError ());
  c->have_prev = true;
  return GFSB200_OK;
}

extern "C" void * gfsb200_host_alloc (size_t bytes)
{
  void * p = NULL;
  if (cudaMallocHost (&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError ();
    gfsb200_fail (GFSB200_ERR_NOMEM, "host_alloc: cannot page-lock %zu bytes", bytes);
    return NULL;
  }
  return p;
}

extern "C" void gfsb200_host_free (void * p)
{
  if (p) cudaFreeHost (p);
}

/* the owned device copies of U,V,W,alpha,mu: allocated for the components in `present', released
 * for the others, and made the resident field */
extern "C" int gfsb200_internal_field_buffers (gfsb200_ctx * c, const int present[5])
{
  const size_t bytes = (size_t) c->T.n_cells*sizeof (double);
  for (int i = 0; i < 5; i++) {
    if (present[i]) {
      /* (+ 16: the bulk staging of lattice_cell_pass_kernel reads one cell past the last leaf) */
      if (!c->d_field[i]) {
	CK (cudaMalloc ((void **) &c->d_field[i], bytes + 16));
	CK (cudaMemsetAsync (c->d_field[i] + c->T.n_cells, 0, 16, c->stream));
      }
    }
    else if (c->d_field[i]) {
      CK (cudaStreamSynchronize (c->stream));
      cudaFree (c->d_field[i]);
      c->d_field[i] = NULL;
    }
  }
  c->F.u[0] = c->d_field[0]; c->F.u[1] = c->d_field[1]; c->F.u[2] = c->d_field[2];
  c->F.alpha = c->d_field[3]; c->F.mu = c->d_field[4];
  return GFSB200_OK;
}

extern "C" int gfsb200_upload_field (gfsb200_ctx * c, const double * u, const double * v,
				     const double * w, const double * alpha, const double * mu)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "upload_field: upload a tree first");
  if (!u || !v || (c->T.dim == 3 && !w))
    return gfsb200_fail (GFSB200_ERR_ARG, "upload_field: missing velocity component");
  CK (cudaSetDevice (c->device));
  const double * src[5] = { u, v, c->T.dim == 3 ? w : NULL, alpha, mu };
  const int present[5] = { 1, 1, c->T.dim == 3, alpha != NULL, mu != NULL };
  int r = gfsb200_internal_field_buffers (c, present);
  if (r) return r;
  const size_t bytes = (size_t) c->T.n_cells*sizeof (double);
  for (int i = 0; i < 5; i++)
    if (present[i])
      CK (cudaMemcpyAsync (c->d_field[i], src[i], bytes, cudaMemcpyHostToDevice, c->stream));
  return gfsb200_refresh_field (c);
}

extern "C" int gfsb200_upload_field_part (gfsb200_ctx * c, int64_t first, int64_t n,
					  const double * u, const double * v, const double * w,
					  const double * alpha, const double * mu)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "upload_field_part: upload a tree first");
  if (!u || !v || (c->T.dim == 3 && !w))
    return gfsb200_fail (GFSB200_ERR_ARG, "upload_field_part: missing velocity component");
  if (first < 0 || n < 0 || first + n > c->T.n_cells)
    return gfsb200_fail (GFSB200_ERR_ARG, "upload_field_part: cells [%lld, %lld) out of range",
			 (long long) first, (long long) (first + n));
  CK (cudaSetDevice (c->device));
  const double * src[5] = { u, v, c->T.dim == 3 ? w : NULL, alpha, mu };
  const int present[5] = { 1, 1, c->T.dim == 3, alpha != NULL, mu != NULL };
  int r = gfsb200_internal_field_buffers (c, present);
  if (r) return r;
  c->have_field = false;                 /* until gfsb200_refresh_field has rebuilt the tables */
  for (int i = 0; i < 5; i++)
    if (present[i] && n)
      CK (cudaMemcpyAsync (c->d_field[i] + first, src[i] + first, (size_t) n*sizeof (double),
			   cudaMemcpyHostToDevice, c->stream));
  return GFSB200_OK;
}

extern "C" int gfsb200_set_field_device (gfsb200_ctx * c, const double * u, const double * v,
					 const double * w, const double * alpha, const double * mu)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "set_field_device: upload a tree first");
  if (!u || !v || (c->T.dim == 3 && !w))
    return gfsb200_fail (GFSB200_ERR_ARG, "set_field_device: missing velocity component");
  c->F.u[0] = u; c->F.u[1] = v; c->F.u[2] = c->T.dim == 3 ? w : NULL;
  c->F.alpha = alpha; c->F.mu = mu;
  return gfsb200_refresh_field (c);
}

extern "C" int gfsb200_download_corner_values (gfsb200_ctx * c, int comp, int64_t n,
					       const int32_t * cells, double * out)
{
  if (!c || !c->have_field) return gfsb200_fail (GFSB200_ERR_STATE, "download_corner_values: no field");
  if (comp < 0 || comp >= c->T.dim || n < 0 || !cells || !out)
    return gfsb200_fail (GFSB200_ERR_ARG, "download_corner_values: bad argument");
  CK (cudaSetDevice (c->device));
  const int nc = 1 << c->T.dim;
  if (n == 0) return GFSB200_OK;
  Scratch sc (c);
  int r = sc.reserve (Scratch::pad (n*sizeof (int32_t)) + Scratch::pad (n*nc*sizeof (double)));
  if (r) return r;
  int32_t * d_cells = sc.take<int32_t> (n);
  double * d_out = sc.take<double> (n*nc);
  CK (cudaMemcpyAsync (d_cells, cells, n*sizeof (int32_t), cudaMemcpyHostToDevice, c->stream));
  gfsb200_launch_corner_values (&c->T, &c->F, comp, n, d_cells, d_out, c->stream);
  CK (cudaGetLastError ());
  CK (cudaMemcpyAsync (out, d_out, n*nc*sizeof (double), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return GFSB200_OK;
}

extern "C" int gfsb200_download_vorticity (gfsb200_ctx * c, int64_t n, const int32_t * cells,
					   double * out)
{
  if (!c || !c->have_field) return gfsb200_fail (GFSB200_ERR_STATE, "download_vorticity: no field");
  if (n < 0 || !cells || !out) return gfsb200_fail (GFSB200_ERR_ARG, "download_vorticity: bad argument");
  CK (cudaSetDevice (c->device));
  const int ws = c->T.dim == 3 ? 4 : 1;
  std::vector<double> all ((size_t) c->T.n_cells*ws);
  CK (cudaMemcpyAsync (all.data (), c->F.vort, all.size ()*sizeof (double), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  for (int64_t j = 0; j < n; j++) {
    if (cells[j] < 0 || cells[j] >= c->T.n_cells)
      return gfsb200_fail (GFSB200_ERR_ARG, "download_vorticity: cell %d out of range", cells[j]);
    size_t slot = cells[j];
    if (c->T.lattice_n1 > 0) {
      const int per_level = 1 << (c->T.dim*c->T.top_levels);
      if (cells[j] < c->T.top_start || cells[j] >= c->T.top_start + per_level)
	return gfsb200_fail (GFSB200_ERR_ARG, "download_vorticity: cell %d is not a leaf", cells[j]);
      slot = gfsb200_lattice_index (c->T.dim, c->T.top_start, c->T.lattice_n1 - 1, cells[j]);
    }
    if (c->T.dim == 3) {           /* split rows: (wx,wy) pairs, then the wz array (DevField) */
      out[3*j] = all[slot*2]; out[3*j + 1] = all[slot*2 + 1]; out[3*j + 2] = all[(size_t) c->T.n_cells*2 + slot];
    }
    else {
      out[3*j] = 0.; out[3*j + 1] = 0.; out[3*j + 2] = all[slot];
    }
  }
  return GFSB200_OK;
}

/* ------------------------------------------------------------------ */
/* particles                                                            */

static int ensure_aux (gfsb200_ctx * c, int64_t n)
{
  if (n <= c->aux_cap) return GFSB200_OK;
  for (int k = 0; k < 3; k++) { cudaFree (c->force[k]); c->force[k] = NULL; }
  cudaFree (c->cell); cudaFree (c->perm); cudaFree (c->perm2); cudaFree (c->key); cudaFree (c->key2);
  cudaFree (c->flag);
  c->aux_cap = 0;
  for (int k = 0; k < 3; k++) CK (cudaMalloc ((void **) &c->force[k], n*sizeof (double)));
  CK (cudaMalloc ((void **) &c->cell, n*sizeof (int32_t)));
  CK (cudaMalloc ((void **) &c->perm, n*sizeof (int32_t)));
  CK (cudaMalloc ((void **) &c->perm2, n*sizeof (int32_t)));
  CK (cudaMalloc ((void **) &c->key, n*sizeof (uint32_t)));
  CK (cudaMalloc ((void **) &c->key2, n*sizeof (uint32_t)));
  CK (cudaMalloc ((void **) &c->flag, n));
  c->aux_cap = n;
  return GFSB200_OK;
}

/* capacity for n particles in both SoA buffers; the current contents (c->n particles) are kept */
extern "C" int gfsb200_internal_reserve (gfsb200_ctx * c, int64_t n)
{
  if (n <= c->cap) return GFSB200_OK;
  CK (cudaStreamSynchronize (c->stream));
  int64_t keep = c->n;
  /* columns are padded to a whole number of 256-particle tiles so that the
     bulk copies of the staged step kernel never run past the allocation */
  const int64_t alloc_n = (n + 255)/256*256;
  double * ncol[2][NCOL]; uint32_t * nid[2];
  for (int b = 0; b < 2; b++) {
    for (int k = 0; k < NCOL; k++) {
      CK (cudaMalloc ((void **) &ncol[b][k], alloc_n*sizeof (double)));
      CK (cudaMemsetAsync (ncol[b][k], 0, alloc_n*sizeof (double), c->stream));
    }
    CK (cudaMalloc ((void **) &nid[b], alloc_n*sizeof (uint32_t)));
  }
  if (keep) {
    for (int k = 0; k < NCOL; k++)
      CK (cudaMemcpyAsync (ncol[0][k], c->col[c->cur][k], keep*sizeof (double), cudaMemcpyDeviceToDevice, c->stream));
    CK (cudaMemcpyAsync (nid[0], c->id[c->cur], keep*sizeof (uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    CK (cudaStreamSynchronize (c->stream));
  }
  for (int b = 0; b < 2; b++) {
    for (int k = 0; k < NCOL; k++) { cudaFree (c->col[b][k]); c->col[b][k] = ncol[b][k]; }
    cudaFree (c->id[b]); c->id[b] = nid[b];
  }
  c->cur = 0;
  c->cap = n;
  double * table[2*NCOL];
  for (int b = 0; b < 2; b++) for (int k = 0; k < NCOL; k++) table[b*NCOL + k] = c->col[b][k];
  CK (cudaMemcpyAsync (c->d_ptr_table, table, sizeof table, cudaMemcpyHostToDevice, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return GFSB200_OK;
}

extern "C" int gfsb200_internal_ensure_aux (gfsb200_ctx * c, int64_t n) { return ensure_aux (c, n); }

extern "C" int gfsb200_particles_resize (gfsb200_ctx * c, int64_t n)
{
  if (!c || n < 0 || n > INT32_MAX)
    return gfsb200_fail (GFSB200_ERR_ARG, "particles_resize: bad count");
  CK (cudaSetDevice (c->device));
  int r = gfsb200_internal_reserve (c, n);
  if (r) return r;
  r = ensure_aux (c, n > 0 ? n : 1);
  if (r) return r;
  if (n > c->n)    /* new slots get ids continuing the sequence */
    gfsb200_launch_iota_u32 (n - c->n, c->id[c->cur] + c->n, (uint32_t) c->n + 1, c->stream);
  c->n = n;
  return GFSB200_OK;
}

extern "C" int gfsb200_particles_device_ptrs (gfsb200_ctx * c, double * ptrs[8])
{
  if (!c || !ptrs) return gfsb200_fail (GFSB200_ERR_ARG, "particles_device_ptrs: null argument");
  for (int k = 0; k < NCOL; k++) ptrs[k] = c->col[c->cur][k];
  return GFSB200_OK;
}

extern "C" int64_t gfsb200_particles_count (gfsb200_ctx * c) { return c ? c->n : -1; }

extern "C" int gfsb200_particles_upload (gfsb200_ctx * c, int64_t n,
					 const double * x, const double * y, const double * z,
					 const double * vx, const double * vy, const double * vz,
					 const double * mass, const double * volume, const uint32_t * id)
{
  if (!c || n < 0 || (n && (!x || !y || !vx || !vy || !mass || !volume)))
    return gfsb200_fail (GFSB200_ERR_ARG, "particles_upload: bad argument");
  c->n = 0;
  int r = gfsb200_particles_resize (c, n);
  if (r) return r;
  const double * src[NCOL] = { x, y, z, vx, vy, vz, mass, volume };
  for (int k = 0; k < NCOL; k++) {
    if (src[k])
      CK (cudaMemcpyAsync (c->col[c->cur][k], src[k], n*sizeof (double), cudaMemcpyHostToDevice, c->stream));
    else
      CK (cudaMemsetAsync (c->col[c->cur][k], 0, n*sizeof (double), c->stream));
  }
  if (id)
    CK (cudaMemcpyAsync (c->id[c->cur], id, n*sizeof (uint32_t), cudaMemcpyHostToDevice, c->stream));
  CK (cudaMemsetAsync (c->cell, 0xff, (n ? n : 1)*sizeof (int32_t), c->stream));
  for (int k = 0; k < 3; k++) CK (cudaMemsetAsync (c->force[k], 0, (n ? n : 1)*sizeof (double), c->stream));
  c->forces_recorded = false;
  CK (cudaStreamSynchronize (c->stream));
  return GFSB200_OK;
}

extern "C" int gfsb200_particles_download (gfsb200_ctx * c,
					   double * x, double * y, double * z,
					   double * vx, double * vy, double * vz,
					   double * fx, double * fy, double * fz,
					   double * mass, double * volume, uint32_t * id, int32_t * cell)
{
  if (!c) return gfsb200_fail (GFSB200_ERR_ARG, "particles_download: null context");
  CK (cudaSetDevice (c->device));
  const int64_t n = c->n;
  double * dst[NCOL] = { x, y, z, vx, vy, vz, mass, volume };
  for (int k = 0; k < NCOL; k++)
    if (dst[k] && n)
      CK (cudaMemcpyAsync (dst[k], c->col[c->cur][k], n*sizeof (double), cudaMemcpyDeviceToHost, c->stream));
  double * fd[3] = { fx, fy, fz };
  for (int k = 0; k < 3; k++)
    if (fd[k] && n)
      CK (cudaMemcpyAsync (fd[k], c->force[k], n*sizeof (double), cudaMemcpyDeviceToHost, c->stream));
  if (id && n) CK (cudaMemcpyAsync (id, c->id[c->cur], n*sizeof (uint32_t), cudaMemcpyDeviceToHost, c->stream));
  if (cell && n) CK (cudaMemcpyAsync (cell, c->cell, n*sizeof (int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return GFSB200_OK;
}

static DevParticles particles_view (gfsb200_ctx * c)
{
  DevParticles P;
  P.n = c->n;
  double ** col = c->col[c->cur];
  P.x = col[0]; P.y = col[1]; P.z = col[2]; P.vx = col[3]; P.vy = col[4]; P.vz = col[5];
  P.mass = col[6]; P.volume = col[7];
  P.fx = c->force[0]; P.fy = c->force[1]; P.fz = c->force[2];
  P.cell = c->cell;
  P.id = c->id[c->cur];
  return P;
}

extern "C" void gfsb200_step_params_default (gfsb200_step_params * p)
{
  if (!p) return;
  memset (p, 0, sizeof *p);
  p->rho = 1.;
  p->cd_const = NAN;
  p->cl_const = NAN;
  p->cm_const = NAN;
}

static int make_step (const gfsb200_step_params * p, DevStep * S)
{
  if (!p) return gfsb200_fail (GFSB200_ERR_ARG, "null step parameters");
  if (p->n_forces < 0 || p->n_forces > GFSB200_MAX_FORCES)
    return gfsb200_fail (GFSB200_ERR_ARG, "n_forces = %d out of range", p->n_forces);
  memset (S, 0, sizeof *S);
  S->dt = p->dt;
  S->n_forces = p->n_forces;
  for (int k = 0; k < p->n_forces; k++) {
    if (p->force[k] < GFSB200_FORCE_DRAG || p->force[k] > GFSB200_FORCE_ADDEDMASS)
      return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "force kind %d is not supported on the device", p->force[k]);
    if (p->force[k] == GFSB200_FORCE_ADDEDMASS) S->mutates_mass = 1;
    S->forces |= (unsigned) p->force[k] << (4*k);
    if (p->force[k] != GFSB200_FORCE_BUOY) S->need_velocity = 1;
  }
  S->rho = p->rho; S->mu = p->mu;
  S->inv_mu = p->mu != 0. ? 1./p->mu : 0.;
  S->inv_rho = 1./p->rho;
  for (int a = 0; a < 3; a++) S->g[a] = p->g[a];
  S->cd_const = p->cd_const; S->cl_const = p->cl_const;
  S->cm_const = p->cm_const;
  return GFSB200_OK;
}

#define MAX_TIMED_EVENTS 16384

/* GfsForceInertial / GfsForceAddedMass need the previous-step vertex table
 * (gfsb200_upload_field_prev) and the per-leaf convective acceleration, which
 * is rebuilt here once per field update, only when such a force is listed */
static int prepare_inertial (gfsb200_ctx * c, const DevStep * S)
{
  bool need = false;
  for (int k = 0; k < S->n_forces; k++) {
    const unsigned kind = (S->forces >> (4*k)) & 15;
    if (kind == GFSB200_FORCE_INERTIAL || kind == GFSB200_FORCE_ADDEDMASS) need = true;
  }
  if (!need) return GFSB200_OK;
  if (!c->have_prev)
    return gfsb200_fail (GFSB200_ERR_STATE, "GfsForceInertial/AddedMass: call gfsb200_upload_field_prev first");
  if (!c->acc_valid) {
    if (!c->F.acc) {
      const int ws = c->T.dim == 3 ? 4 : 2;
      CK (cudaMalloc ((void **) &c->F.acc, (size_t) c->T.n_cells*ws*sizeof (double)));
      CK (cudaMemsetAsync (c->F.acc, 0, (size_t) c->T.n_cells*ws*sizeof (double), c->stream));
    }
    gfsb200_launch_convective (&c->T, &c->F, c->n_sm, c->stream);
    CK (cudaGetLastError ());
    c->acc_valid = true;
  }
  return GFSB200_OK;
}

static int timed_begin (gfsb200_ctx * c)
{
  c->timing = c->ev_used + 2 <= MAX_TIMED_EVENTS;   /* stop recording until the next timer_reset */
  /* gfsb200_timer_sampling: the two events of a timed launch sit between the cell pass and the step kernel
     in the stream (and keep the latter's programmatic launch from overlapping the former) -- 1.5 % of a
     C2 step when every launch carries them */
  if (c->timer_every <= 0 || (c->timer_every > 1 && (c->timer_calls++ % c->timer_every) != 0)) c->timing = false;
  if (!c->timing) return GFSB200_OK;
  if (c->ev_used + 2 > c->ev.size ()) {
    cudaEvent_t a, b;
    CK (cudaEventCreate (&a));
    CK (cudaEventCreate (&b));
    c->ev.push_back (a); c->ev.push_back (b);
  }
  CK (cudaEventRecord (c->ev[c->ev_used], c->stream));
  return GFSB200_OK;
}

static int timed_end (gfsb200_ctx * c)
{
  if (!c->timing) return GFSB200_OK;
  CK (cudaEventRecord (c->ev[c->ev_used + 1], c->stream));
  c->ev_used += 2;
  return GFSB200_OK;
}

static int deposit_target (gfsb200_ctx * c, int what, bool local_only, DevDeposit * D);

extern "C" int gfsb200_step (gfsb200_ctx * c, const gfsb200_step_params * p)
{
  if (!c || !c->have_field)
    return gfsb200_fail (GFSB200_ERR_STATE, "step: tree and field must be resident");
  DevStep S;
  int r = make_step (p, &S);
  if (r) return r;
  CK (cudaSetDevice (c->device));
  if ((r = prepare_inertial (c, &S))) return r;
  c->esc_armed = false;
  if (p->track_escapes && S.n_forces > 0) {
    /* room for 1/16 of the list (+1024) to leave the domain in one step */
    const int want = (int) (c->n/16 + 1024);
    if (want > c->esc_cap) {
      cudaFree (c->esc_idx); cudaFree (c->esc_old);
      c->esc_idx = NULL; c->esc_old = NULL; c->esc_cap = 0;
      CK (cudaMalloc ((void **) &c->esc_idx, want*sizeof (int32_t)));
      CK (cudaMalloc ((void **) &c->esc_old, (size_t) want*3*sizeof (double)));
      c->esc_cap = want;
    }
    if (!c->esc_count) CK (cudaMalloc ((void **) &c->esc_count, 4*sizeof (int)));
    CK (cudaMemsetAsync (c->esc_count, 0, 4*sizeof (int), c->stream));
    S.track_escapes = 1;
    S.esc_cap = c->esc_cap; S.esc_count = c->esc_count; S.esc_idx = c->esc_idx; S.esc_old = c->esc_old;
    S.keep = NULL;
    if (c->mark_outside) {
      /* list event: the particles the kernel finds outside before the step are flagged, so that
	 remove_particles_not_in_domain becomes a compaction after the step, and only if needed */
      CK (cudaMemsetAsync (c->flag, 1, (size_t) (c->n ? c->n : 1), c->stream));
      S.keep = c->flag;
    }
    c->esc_armed = true;
  }
  DevParticles P = particles_view (c);
  DevDeposit D;
  const bool fuse = p->fuse_deposit && S.n_forces > 0;
  if (p->fuse_deposit && !fuse)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "step: fuse_deposit needs a force list (tracers deposit no force)");
  if (fuse && (r = deposit_target (c, 3, false, &D))) return r;
  if ((r = timed_begin (c))) return r;
  int fused = 0;
  if (S.n_forces == 0)
    gfsb200_launch_advect (&c->T, &c->F, &P, S.dt, p->record_cells, c->stream);
  else
    fused = gfsb200_launch_step (&c->T, &c->F, &P, &S, p->record_cells || p->record_forces, c->step_minb,
				 c->step_mode, c->n_sm, c->stream, fuse ? &D : NULL);
  if (fuse && !fused) {
    /* a kernel flavour without the fused tail (runtime force list, recorded forces): the same
       result from the stand-alone pass over the new state */
    gfsb200_launch_deposit (&c->T, &c->F, &P, &S, 3, &D, 0, NULL, NULL, c->stream);
  }
  if ((r = timed_end (c))) return r;
  CK (cudaGetLastError ());
  if (S.n_forces > 0 && p->record_forces)
    c->forces_recorded = true;
  c->last_step_fused = fuse;
  if (fuse) { c->last_S = S; c->last_D = D; }
  return GFSB200_OK;
}

/* ------------------------------------------------------------------ */
/* host-resident particle lists: chunked, triple-buffered H2D -> step -> D2H */

#define HOST_SLOTS 3

static int ensure_host_pipeline (gfsb200_ctx * c, int64_t chunk)
{
  const int64_t alloc = (chunk + 255)/256*256;
  if (c->hp_chunk >= alloc) return GFSB200_OK;
  for (int s = 0; s < HOST_SLOTS; s++)
    for (int k = 0; k < NCOL; k++) {
      cudaFree (c->hp_col[s][k]);
      c->hp_col[s][k] = NULL;
    }
  c->hp_chunk = 0;
  for (int s = 0; s < HOST_SLOTS; s++)
    for (int k = 0; k < NCOL; k++) {
      CK (cudaMalloc ((void **) &c->hp_col[s][k], alloc*sizeof (double)));
      CK (cudaMemsetAsync (c->hp_col[s][k], 0, alloc*sizeof (double), c->stream));
    }
  if (!c->hp_h2d) {
    CK (cudaStreamCreateWithFlags (&c->hp_h2d, cudaStreamNonBlocking));
    CK (cudaStreamCreateWithFlags (&c->hp_d2h, cudaStreamNonBlocking));
    for (int s = 0; s < HOST_SLOTS; s++) {
      CK (cudaEventCreateWithFlags (&c->hp_in[s], cudaEventDisableTiming));
      CK (cudaEventCreateWithFlags (&c->hp_done[s], cudaEventDisableTiming));
      CK (cudaEventCreateWithFlags (&c->hp_out[s], cudaEventDisableTiming));
    }
  }
  CK (cudaStreamSynchronize (c->stream));
  c->hp_chunk = alloc;
  return GFSB200_OK;
}

extern "C" int gfsb200_step_host (gfsb200_ctx * c, const gfsb200_step_params * p, int64_t n,
				  double * x, double * y, double * z,
				  double * vx, double * vy, double * vz,
				  const double * mass, const double * volume, int64_t chunk)
{
  if (!c || !c->have_field)
    return gfsb200_fail (GFSB200_ERR_STATE, "step_host: tree and field must be resident");
  if (n < 0 || (n && (!x || !y || !vx || !vy || !mass || !volume || (c->T.dim == 3 && (!z || !vz)))))
    return gfsb200_fail (GFSB200_ERR_ARG, "step_host: bad argument");
  DevStep S;
  int r = make_step (p, &S);
  if (r) return r;
  if (S.n_forces == 0)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "step_host: tracer lists go through gfsb200_step");
  if (n == 0) return GFSB200_OK;
  CK (cudaSetDevice (c->device));
  if (p->fuse_deposit)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "step_host: fuse_deposit needs the resident list");
  if (S.mutates_mass)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "step_host: GfsForceAddedMass rewrites mass; use the resident path");
  if ((r = prepare_inertial (c, &S))) return r;
  if (chunk <= 0) chunk = 1 << 20;
  if (chunk > n) chunk = n;
  if ((r = ensure_host_pipeline (c, chunk))) return r;

  const bool d3 = c->T.dim == 3;
  const double * in[NCOL] = { x, y, z, vx, vy, vz, mass, volume };
  double * out[6] = { x, y, z, vx, vy, vz };
  const int64_t n_chunks = (n + chunk - 1)/chunk;
  for (int64_t k = 0; k < n_chunks; k++) {
    const int s = (int) (k % HOST_SLOTS);
    const int64_t lo = k*chunk, m = (lo + chunk <= n ? chunk : n - lo);
    /* slot s is free once the D2H of chunk k - HOST_SLOTS has drained */
    if (k >= HOST_SLOTS)
      CK (cudaStreamWaitEvent (c->hp_h2d, c->hp_out[s], 0));
    for (int q = 0; q < NCOL; q++)
      if (in[q] && (d3 || (q != 2 && q != 5)))
	CK (cudaMemcpyAsync (c->hp_col[s][q], in[q] + lo, m*sizeof (double), cudaMemcpyHostToDevice, c->hp_h2d));
    CK (cudaEventRecord (c->hp_in[s], c->hp_h2d));
    CK (cudaStreamWaitEvent (c->stream, c->hp_in[s], 0));
    DevParticles P;
    memset (&P, 0, sizeof P);
    P.n = m;
    P.x = c->hp_col[s][0]; P.y = c->hp_col[s][1]; P.z = c->hp_col[s][2];
    P.vx = c->hp_col[s][3]; P.vy = c->hp_col[s][4]; P.vz = c->hp_col[s][5];
    P.mass = c->hp_col[s][6]; P.volume = c->hp_col[s][7];
    gfsb200_launch_step (&c->T, &c->F, &P, &S, 0, c->step_minb, c->step_mode, c->n_sm, c->stream, NULL);
    CK (cudaGetLastError ());
    CK (cudaEventRecord (c->hp_done[s], c->stream));
    CK (cudaStreamWaitEvent (c->hp_d2h, c->hp_done[s], 0));
    for (int q = 0; q < 6; q++)
      if (out[q] && (d3 || (q != 2 && q != 5)))
	CK (cudaMemcpyAsync (out[q] + lo, c->hp_col[s][q], m*sizeof (double), cudaMemcpyDeviceToHost, c->hp_d2h));
    CK (cudaEventRecord (c->hp_out[s], c->hp_d2h));
  }
  CK (cudaStreamSynchronize (c->hp_d2h));
  CK (cudaStreamSynchronize (c->stream));
  return GFSB200_OK;
}

static int ensure_cub_tmp (gfsb200_ctx * c, size_t bytes)
{
  if (bytes <= c->cub_tmp_bytes) return GFSB200_OK;
  cudaFree (c->cub_tmp);
  c->cub_tmp = NULL; c->cub_tmp_bytes = 0;
  CK (cudaMalloc (&c->cub_tmp, bytes));
  c->cub_tmp_bytes = bytes;
  return GFSB200_OK;
}

/* apply c->perm2 (new position -> old position) to the SoA, n_new entries */
static int apply_permutation (gfsb200_ctx * c, int64_t n_new)
{
  const int nb = 1 - c->cur, old = c->cur;
  gfsb200_launch_gather (n_new, c->perm2, NCOL, (const double * const *) (c->d_ptr_table + c->cur*NCOL),
			 c->d_ptr_table + nb*NCOL, c->id[c->cur], c->id[nb], c->stream);
  CK (cudaGetLastError ());
  c->cur = nb;
  c->n = n_new;
  if (c->forces_recorded && n_new > 0) {
    /* the forces recorded by the last step follow their particles; the columns just vacated
       serve as scratch */
    gfsb200_launch_gather3 (n_new, c->perm2, c->force[0], c->force[1], c->force[2],
			    c->col[old][0], c->col[old][1], c->col[old][2], c->stream);
    CK (cudaGetLastError ());
    for (int k = 0; k < 3; k++)
      CK (cudaMemcpyAsync (c->force[k], c->col[old][k], n_new*sizeof (double), cudaMemcpyDeviceToDevice,
			   c->stream));
  }
  return GFSB200_OK;
}

/* sort the resident particles by containing cell; the sorted keys (flat cell index, n_cells for a
 * particle outside the domain) are left in c->key2 */
extern "C" int gfsb200_internal_sort (gfsb200_ctx * c)
{
  if (c->n <= 0) return GFSB200_OK;
  DevParticles P = particles_view (c);
  gfsb200_launch_locate (&c->T, P.n, P.x, P.y, P.z, c->cell, c->stream);
  gfsb200_launch_sort_keys (P.n, c->cell, c->key, (uint32_t) c->T.n_cells, c->stream);
  gfsb200_launch_iota (P.n, c->perm, c->stream);
  CK (cudaGetLastError ());
  int end_bit = 1;
  while (end_bit < 32 && (1u << end_bit) <= (uint32_t) c->T.n_cells) end_bit++;
  size_t bytes = 0;
  CK (gfsb200_cub_sort_pairs (NULL, &bytes, c->key, c->key2, c->perm, c->perm2, P.n, end_bit, c->stream));
  int r = ensure_cub_tmp (c, bytes);
  if (r) return r;
  CK (gfsb200_cub_sort_pairs (c->cub_tmp, &bytes, c->key, c->key2, c->perm, c->perm2, P.n, end_bit, c->stream));
  return apply_permutation (c, P.n);
}

/* Right after gfsb200_internal_sort: a STABLE re-sort of the cell-sorted list by the rank that owns each
 * particle's cell (owner_of, device memory; particles outside the domain last), so that every rank's
 * share is one contiguous stretch, still sorted by cell inside.  The sorted owner keys (nranks for the
 * particles outside) are left in c->key2. */
extern "C" int gfsb200_internal_sort_by_owner (gfsb200_ctx * c, const uint8_t * owner_of, int nranks)
{
  if (c->n <= 0) return GFSB200_OK;
  const int64_t n = c->n;
  gfsb200_launch_owner_keys (n, c->key2, (uint32_t) c->T.n_cells, owner_of, c->key, (uint32_t) nranks, c->stream);
  gfsb200_launch_iota (n, c->perm, c->stream);
  CK (cudaGetLastError ());
  int end_bit = 1;
  while (end_bit < 32 && (1u << end_bit) <= (uint32_t) nranks) end_bit++;
  size_t bytes = 0;
  CK (gfsb200_cub_sort_pairs (NULL, &bytes, c->key, c->key2, c->perm, c->perm2, n, end_bit, c->stream));
  int r = ensure_cub_tmp (c, bytes);
  if (r) return r;
  CK (gfsb200_cub_sort_pairs (c->cub_tmp, &bytes, c->key, c->key2, c->perm, c->perm2, n, end_bit, c->stream));
  return apply_permutation (c, n);
}

extern "C" int gfsb200_particles_sort (gfsb200_ctx * c)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "particles_sort: no tree");
  if (c->n <= 1) return GFSB200_OK;
  CK (cudaSetDevice (c->device));
  return gfsb200_internal_sort (c);
}

extern "C" int gfsb200_particles_cull (gfsb200_ctx * c, int64_t * n_removed)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "particles_cull: no tree");
  if (n_removed) *n_removed = 0;
  if (c->n == 0) return GFSB200_OK;
  CK (cudaSetDevice (c->device));
  DevParticles P = particles_view (c);
  gfsb200_launch_locate (&c->T, P.n, P.x, P.y, P.z, c->cell, c->stream);
  gfsb200_launch_inside_flags (P.n, c->cell, c->flag, c->stream);
  gfsb200_launch_iota (P.n, c->perm, c->stream);
  CK (cudaGetLastError ());
  size_t bytes = 0;
  CK (gfsb200_cub_select_flagged (NULL, &bytes, c->perm, c->flag, c->perm2, c->d_count, P.n, c->stream));
  int r = ensure_cub_tmp (c, bytes);
  if (r) return r;
  CK (gfsb200_cub_select_flagged (c->cub_tmp, &bytes, c->perm, c->flag, c->perm2, c->d_count, P.n, c->stream));
  int32_t kept = 0;
  CK (cudaMemcpyAsync (&kept, c->d_count, sizeof kept, cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  if (kept == P.n) return GFSB200_OK;
  if (n_removed) *n_removed = P.n - kept;
  return apply_permutation (c, kept);
}

/* flags_ready: c->flag already holds 1 / 0 (kept / outside before the step, `outside` of them);
 * the particles dropped here are cleared in the same array and ONE compaction removes both */
static int particle_bc_impl (gfsb200_ctx * c, bool flags_ready, int outside, int64_t * n_wrapped,
			     int64_t * n_dropped)
{
  if (n_wrapped) *n_wrapped = 0;
  if (n_dropped) *n_dropped = 0;
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "particle_bc: no tree");
  if (!c->esc_armed)
    return gfsb200_fail (GFSB200_ERR_STATE, "particle_bc: the last step did not track escapes");
  CK (cudaSetDevice (c->device));
  c->esc_armed = false;
  int counts[3] = { 0, 0, 0 };
  CK (cudaMemcpyAsync (counts, c->esc_count, sizeof (int), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  if (counts[0] == 0 && outside == 0) return GFSB200_OK;
  /* More particles left in one step than the record holds (n/16 + 1024): the first esc_cap are
     handled exactly; the others have lost their previous position (the step has overwritten it),
     so the exit face cannot be found -- they are dropped, as a particle leaving through a
     non-periodic side is (the pass below removes whatever is still outside). */
  const bool overflow = counts[0] > c->esc_cap;
  const int n_esc = overflow ? c->esc_cap : counts[0];
  DevParticles P = particles_view (c);
  if (!flags_ready)
    CK (cudaMemsetAsync (c->flag, 1, (size_t) P.n, c->stream));
  if (n_esc > 0) {
    gfsb200_launch_particle_bc (&c->T, &P, n_esc, c->esc_idx, c->esc_old, c->flag, c->esc_count + 1,
				c->stream);
    CK (cudaGetLastError ());
    if (c->last_step_fused) {
      /* the fused step deposited nothing for a particle whose new position was outside; those
	 that were wrapped back in deposit now, at the wrapped position */
      gfsb200_launch_deposit (&c->T, &c->F, &P, &c->last_S, 3, &c->last_D, n_esc, c->esc_idx, c->flag, c->stream);
      CK (cudaGetLastError ());
    }
    CK (cudaMemcpyAsync (counts, c->esc_count, 3*sizeof (int), cudaMemcpyDeviceToHost, c->stream));
    CK (cudaStreamSynchronize (c->stream));
  }
  int lost = 0;
  if (overflow) {
    gfsb200_launch_locate (&c->T, P.n, P.x, P.y, P.z, c->cell, c->stream);
    gfsb200_launch_outside_clear (P.n, c->cell, c->flag, c->esc_count + 1, c->stream);
    CK (cudaGetLastError ());
    int w = counts[1];
    CK (cudaMemcpyAsync (&lost, c->esc_count + 1, sizeof (int), cudaMemcpyDeviceToHost, c->stream));
    CK (cudaStreamSynchronize (c->stream));
    lost -= w;                 /* esc_count[1] was reused as the counter of the clearing pass */
    counts[2] += lost;
  }
  if (n_wrapped) *n_wrapped = counts[1];
  if (n_dropped) *n_dropped = counts[2];
  counts[2] += outside;
  if (counts[2] == 0) return GFSB200_OK;
  /* compact the list without the dropped particles (order kept) */
  gfsb200_launch_iota (P.n, c->perm, c->stream);
  size_t bytes = 0;
  CK (gfsb200_cub_select_flagged (NULL, &bytes, c->perm, c->flag, c->perm2, c->d_count, P.n, c->stream));
  int r = ensure_cub_tmp (c, bytes);
  if (r) return r;
  CK (gfsb200_cub_select_flagged (c->cub_tmp, &bytes, c->perm, c->flag, c->perm2, c->d_count, P.n, c->stream));
  return apply_permutation (c, P.n - counts[2]);
}

extern "C" int gfsb200_particle_bc (gfsb200_ctx * c, int64_t * n_wrapped, int64_t * n_dropped)
{
  return particle_bc_impl (c, false, 0, n_wrapped, n_dropped);
}

extern "C" int gfsb200_escaped_count (gfsb200_ctx * c, int64_t * n_escaped)
{
  if (!c || !n_escaped) return gfsb200_fail (GFSB200_ERR_ARG, "escaped_count: bad argument");
  *n_escaped = 0;
  if (!c->esc_armed)
    return gfsb200_fail (GFSB200_ERR_STATE, "escaped_count: the last step did not track escapes");
  CK (cudaSetDevice (c->device));
  int n = 0;
  CK (cudaMemcpyAsync (&n, c->esc_count, sizeof (int), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  *n_escaped = n;
  return GFSB200_OK;
}

extern "C" int gfsb200_step_counts (gfsb200_ctx * c, int64_t * n_escaped, int64_t * n_outside)
{
  if (!c) return gfsb200_fail (GFSB200_ERR_ARG, "step_counts: null context");
  if (n_escaped) *n_escaped = 0;
  if (n_outside) *n_outside = 0;
  if (!c->esc_armed)
    return gfsb200_fail (GFSB200_ERR_STATE, "step_counts: the last step did not track escapes");
  CK (cudaSetDevice (c->device));
  int counts[4] = { 0, 0, 0, 0 };
  CK (cudaMemcpyAsync (counts, c->esc_count, 4*sizeof (int), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  if (n_escaped) *n_escaped = counts[0];
  if (n_outside) *n_outside = counts[3];
  return GFSB200_OK;
}

extern "C" int gfsb200_escaped_download (gfsb200_ctx * c, int64_t cap, int32_t * idx, double * old_xyz,
					  int64_t * n_out)
{
  if (!c || cap < 0 || !n_out || (cap && (!idx || !old_xyz)))
    return gfsb200_fail (GFSB200_ERR_ARG, "escaped_download: bad argument");
  int64_t n = 0;
  int r = gfsb200_escaped_count (c, &n);
  if (r) return r;
  if (n > c->esc_cap) n = c->esc_cap;   /* the record holds n/16 + 1024 particles; *n_out tells the caller */
  if (n > cap) n = cap;
  if (n) {
    CK (cudaMemcpyAsync (idx, c->esc_idx, n*sizeof (int32_t), cudaMemcpyDeviceToHost, c->stream));
    CK (cudaMemcpyAsync (old_xyz, c->esc_old, 3*n*sizeof (double), cudaMemcpyDeviceToHost, c->stream));
    CK (cudaStreamSynchronize (c->stream));
  }
  *n_out = n;
  return GFSB200_OK;
}

extern "C" int gfsb200_particle_list_event (gfsb200_ctx * c, const gfsb200_step_params * p,
					    int64_t * n_removed)
{
  if (!p) return gfsb200_fail (GFSB200_ERR_ARG, "null step parameters");
  int64_t culled = 0, dropped = 0;
  int r;
  gfsb200_step_params q = *p;
  if (q.n_forces == 0) {
    /* tracers: the RK2 kernel does not track, cull first as the reference does */
    if ((r = gfsb200_particles_cull (c, &culled))) return r;
    q.track_escapes = 0;
    if ((r = gfsb200_step (c, &q))) return r;
    if (n_removed) *n_removed = culled;
    return GFSB200_OK;
  }
  /* The step kernel leaves a particle that is outside the domain untouched; here it also flags
     and counts it, and it counts those that leave.  remove_particles_not_in_domain then is a
     compaction AFTER the step, shared with the particles gfs_particle_bc drops -- the same list
     as cull -> step -> BC -- and neither the cull pass (locate + flag + select, about as long as
     the step itself) nor the BC pass runs when its count is zero: one 16-byte read-back decides. */
  q.track_escapes = 1;
  c->mark_outside = true;
  r = gfsb200_step (c, &q);
  c->mark_outside = false;
  if (r) return r;
  int counts[4] = { 0, 0, 0, 0 };
  CK (cudaMemcpyAsync (counts, c->esc_count, 4*sizeof (int), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  culled = counts[3];
  if (counts[0] > 0 || culled > 0) {
    if ((r = particle_bc_impl (c, true, (int) culled, NULL, &dropped))) return r;
  }
  else
    c->esc_armed = false;
  if (n_removed) *n_removed = culled + dropped;
  return GFSB200_OK;
}

/* ------------------------------------------------------------------ */
/* batched point queries                                                */

extern "C" int gfsb200_locate (gfsb200_ctx * c, int64_t n, const double * x, const double * y,
			       const double * z, int32_t * cell)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "locate: no tree resident");
  if (n < 0 || (n && (!x || !y || !cell || (c->T.dim == 3 && !z))))
    return gfsb200_fail (GFSB200_ERR_ARG, "locate: bad argument");
  if (n == 0) return GFSB200_OK;
  CK (cudaSetDevice (c->device));
  double * d[3] = { NULL, NULL, NULL }; int32_t * dc = NULL;
  const double * src[3] = { x, y, z };
  Scratch sc (c);
  int r = sc.reserve (3*Scratch::pad (n*sizeof (double)) + Scratch::pad (n*sizeof (int32_t)));
  if (r) return r;
  for (int a = 0; a < c->T.dim; a++) {
    d[a] = sc.take<double> (n);
    CK (cudaMemcpyAsync (d[a], src[a], n*sizeof (double), cudaMemcpyHostToDevice, c->stream));
  }
  dc = sc.take<int32_t> (n);
  gfsb200_launch_locate (&c->T, n, d[0], d[1], d[2], dc, c->stream);
  CK (cudaGetLastError ());
  CK (cudaMemcpyAsync (cell, dc, n*sizeof (int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return GFSB200_OK;
}

extern "C" int gfsb200_interpolate (gfsb200_ctx * c, int64_t n, const double * x, const double * y,
				    const double * z, double * u, double * v, double * w)
{
  if (!c || !c->have_field) return gfsb200_fail (GFSB200_ERR_STATE, "interpolate: no field resident");
  if (n < 0 || (n && (!x || !y || (c->T.dim == 3 && !z))))
    return gfsb200_fail (GFSB200_ERR_ARG, "interpolate: bad argument");
  if (n == 0) return GFSB200_OK;
  CK (cudaSetDevice (c->device));
  double * d[3] = { NULL, NULL, NULL }, * o[3] = { NULL, NULL, NULL };
  const double * src[3] = { x, y, z };
  double * dst[3] = { u, v, c->T.dim == 3 ? w : NULL };
  Scratch sc (c);
  int r = sc.reserve (6*Scratch::pad (n*sizeof (double)));
  if (r) return r;
  for (int a = 0; a < c->T.dim; a++) {
    d[a] = sc.take<double> (n);
    CK (cudaMemcpyAsync (d[a], src[a], n*sizeof (double), cudaMemcpyHostToDevice, c->stream));
    if (dst[a]) o[a] = sc.take<double> (n);
  }
  gfsb200_launch_interpolate (&c->T, &c->F, n, d[0], d[1], d[2], o[0], o[1], o[2], c->stream);
  CK (cudaGetLastError ());
  for (int a = 0; a < c->T.dim; a++)
    if (dst[a])
      CK (cudaMemcpyAsync (dst[a], o[a], n*sizeof (double), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return GFSB200_OK;
}

/* GfsOutputLocation for arbitrary cell variables, three per pass (the vertex kernel and
 * the interpolation kernel work on component triples) */
extern "C" int gfsb200_output_location (gfsb200_ctx * c, int nvar, const double * const * vars,
					int interpolate, int64_t n, const double * x, const double * y,
					const double * z, double * const * out, int32_t * cell)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "output_location: no tree resident");
  if (nvar < 0 || n < 0 || (nvar && (!vars || !out)) || (n && (!x || !y || (c->T.dim == 3 && !z))))
    return gfsb200_fail (GFSB200_ERR_ARG, "output_location: bad argument");
  for (int k = 0; k < nvar; k++)
    if (!vars[k] || !out[k]) return gfsb200_fail (GFSB200_ERR_ARG, "output_location: null variable %d", k);
  if (n == 0) return GFSB200_OK;
  CK (cudaSetDevice (c->device));
  const int dim = c->T.dim;
  const size_t nc = c->T.n_cells;
  double * d[3] = { NULL, NULL, NULL }, * o[3] = { NULL, NULL, NULL }, * f[3] = { NULL, NULL, NULL };
  double * vtx = NULL;
  int * flag = NULL;
  int32_t * dcell = NULL;
  std::vector<int32_t> hcell (n);
  const int vs = dim == 3 ? 4 : 2;
  const size_t nv = (size_t) (c->T.n_vertices ? c->T.n_vertices : 1);
  Scratch sc (c);
  int rc = sc.reserve (6*Scratch::pad (n*sizeof (double)) + Scratch::pad (n*sizeof (int32_t)) +
		       (interpolate ? 3*Scratch::pad (nc*sizeof (double)) + Scratch::pad (nv*vs*sizeof (double)) : 0) + 256);
  if (rc) return rc;
  const double * src[3] = { x, y, z };
  for (int a = 0; a < dim; a++) {
    d[a] = sc.take<double> (n);
    CK (cudaMemcpyAsync (d[a], src[a], n*sizeof (double), cudaMemcpyHostToDevice, c->stream));
  }
  dcell = sc.take<int32_t> (n);
  gfsb200_launch_locate (&c->T, n, d[0], d[1], d[2], dcell, c->stream);
  CK (cudaGetLastError ());
  CK (cudaMemcpyAsync (hcell.data (), dcell, n*sizeof (int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  if (cell) memcpy (cell, hcell.data (), n*sizeof (int32_t));
  if (!interpolate) {
    /* GFS_VALUE (cell, v) */
    for (int k = 0; k < nvar; k++)
      for (int64_t i = 0; i < n; i++)
	out[k][i] = hcell[i] >= 0 ? vars[k][hcell[i]] : GFSB200_NODATA;
    return GFSB200_OK;
  }
  for (int a = 0; a < dim; a++) {
    f[a] = sc.take<double> (nc);
    o[a] = sc.take<double> (n);
  }
  vtx = sc.take<double> (nv*vs);
  flag = sc.take<int> (1);
  for (int k0 = 0; k0 < nvar; k0 += dim) {
    DevField tmp;
    memset (&tmp, 0, sizeof tmp);
    for (int a = 0; a < dim; a++) {
      const int k = k0 + a < nvar ? k0 + a : k0;      /* pad a short batch with its first variable */
      CK (cudaMemcpyAsync (f[a], vars[k], nc*sizeof (double), cudaMemcpyHostToDevice, c->stream));
      tmp.u[a] = f[a];
    }
    tmp.vtx_val = vtx;
    tmp.nodata_flag = flag;
    tmp.nodata_epoch = 1;
    CK (cudaMemsetAsync (flag, 0, sizeof (int), c->stream));
    gfsb200_launch_vertex_values (&c->T, &tmp, c->n_sm, c->stream);
    CK (cudaGetLastError ());
    gfsb200_launch_interpolate (&c->T, &tmp, n, d[0], d[1], dim == 3 ? d[2] : NULL, o[0], o[1], dim == 3 ? o[2] : NULL, c->stream);
    CK (cudaGetLastError ());
    for (int a = 0; a < dim && k0 + a < nvar; a++)
      CK (cudaMemcpyAsync (out[k0 + a], o[a], n*sizeof (double), cudaMemcpyDeviceToHost, c->stream));
    CK (cudaStreamSynchronize (c->stream));
  }
  return GFSB200_OK;
}

/* ------------------------------------------------------------------ */
/* particle data formats                                                */

extern "C" int gfsb200_particles_write_gfs (gfsb200_ctx * c, const char * path, const char * class_name,
					    double L, int append)
{
  if (!c || !path || !class_name) return gfsb200_fail (GFSB200_ERR_ARG, "particles_write_gfs: null argument");
  const int64_t n = c->n;
  const int dim = c->have_tree ? c->T.dim : 3;
  std::vector<double> col[11];
  std::vector<uint32_t> id (n ? n : 1);
  for (int k = 0; k < 11; k++) col[k].assign (n ? n : 1, 0.);
  int r = gfsb200_particles_download (c, col[0].data (), col[1].data (), dim == 3 ? col[2].data () : NULL,
				      col[3].data (), col[4].data (), dim == 3 ? col[5].data () : NULL,
				      c->force[0] ? col[8].data () : NULL, c->force[1] ? col[9].data () : NULL,
				      c->force[2] && dim == 3 ? col[10].data () : NULL,
				      col[6].data (), col[7].data (), id.data (), NULL);
  if (r) return r;
  FILE * fp = fopen (path, append ? "a" : "w");
  if (!fp) return gfsb200_fail (GFSB200_ERR_ARG, "particles_write_gfs: cannot open '%s'", path);
  const double Ld = pow (L, dim);
  for (int64_t i = 0; i < n; i++) {
    /* gfs_event_list_write indent; gfs_particle_write; gfs_particulate_write (two fprintf calls) */
    fputs ("    ", fp);
    fprintf (fp, "%s", class_name);
    fprintf (fp, " %d %g %g %g", (int) id[i], col[0][i], col[1][i], col[2][i]);
    fprintf (fp, " %g %g %g %g %g", col[6][i], col[7][i]*Ld, col[3][i], col[4][i], col[5][i]);
    fprintf (fp, " %g %g %g", col[8][i], col[9][i], col[10][i]);
    fputc ('\n', fp);
  }
  if (fclose (fp) != 0) return gfsb200_fail (GFSB200_ERR_ARG, "particles_write_gfs: write to '%s' failed", path);
  return GFSB200_OK;
}

struct CheckpointHeader {
  char magic[8];
  uint32_t version, dim;
  int64_t n;
  uint32_t has_force, reserved;
};

extern "C" int gfsb200_checkpoint_save (gfsb200_ctx * c, const char * path)
{
  if (!c || !path) return gfsb200_fail (GFSB200_ERR_ARG, "checkpoint_save: null argument");
  const int64_t n = c->n;
  CheckpointHeader h;
  memset (&h, 0, sizeof h);
  memcpy (h.magic, "GFSB200P", 8);
  h.version = 1; h.dim = c->have_tree ? c->T.dim : 3; h.n = n;
  h.has_force = c->force[0] != NULL && c->aux_cap >= n;
  std::vector<double> col[11];
  std::vector<uint32_t> id (n ? n : 1);
  for (int k = 0; k < 11; k++) col[k].assign (n ? n : 1, 0.);
  int r = gfsb200_particles_download (c, col[0].data (), col[1].data (), col[2].data (), col[3].data (),
				      col[4].data (), col[5].data (),
				      h.has_force ? col[8].data () : NULL, h.has_force ? col[9].data () : NULL,
				      h.has_force ? col[10].data () : NULL,
				      col[6].data (), col[7].data (), id.data (), NULL);
  if (r) return r;
  FILE * fp = fopen (path, "wb");
  if (!fp) return gfsb200_fail (GFSB200_ERR_ARG, "checkpoint_save: cannot open '%s'", path);
  bool ok = fwrite (&h, sizeof h, 1, fp) == 1;
  for (int k = 0; k < 11 && ok; k++)
    ok = n == 0 || fwrite (col[k].data (), sizeof (double), n, fp) == (size_t) n;
  ok = ok && (n == 0 || fwrite (id.data (), sizeof (uint32_t), n, fp) == (size_t) n);
  ok = (fclose (fp) == 0) && ok;
  return ok ? GFSB200_OK : gfsb200_fail (GFSB200_ERR_ARG, "checkpoint_save: write to '%s' failed", path);
}

extern "C" int gfsb200_checkpoint_load (gfsb200_ctx * c, const char * path)
{
  if (!c || !path) return gfsb200_fail (GFSB200_ERR_ARG, "checkpoint_load: null argument");
  FILE * fp = fopen (path, "rb");
  if (!fp) return gfsb200_fail (GFSB200_ERR_ARG, "checkpoint_load: cannot open '%s'", path);
  CheckpointHeader h;
  if (fread (&h, sizeof h, 1, fp) != 1 || memcmp (h.magic, "GFSB200P", 8) || h.version != 1 || h.n < 0 ||
      h.n > INT32_MAX) {
    fclose (fp);
    return gfsb200_fail (GFSB200_ERR_ARG, "checkpoint_load: '%s' is not a version-1 GFSB200P checkpoint", path);
  }
  if (c->have_tree && (int) h.dim != c->T.dim) {
    fclose (fp);
    return gfsb200_fail (GFSB200_ERR_ARG, "checkpoint_load: checkpoint is %uD, the resident tree %dD", h.dim, c->T.dim);
  }
  const int64_t n = h.n;
  std::vector<double> col[11];
  std::vector<uint32_t> id (n ? n : 1);
  bool ok = true;
  for (int k = 0; k < 11 && ok; k++) {
    col[k].assign (n ? n : 1, 0.);
    ok = n == 0 || fread (col[k].data (), sizeof (double), n, fp) == (size_t) n;
  }
  ok = ok && (n == 0 || fread (id.data (), sizeof (uint32_t), n, fp) == (size_t) n);
  fclose (fp);
  if (!ok) return gfsb200_fail (GFSB200_ERR_ARG, "checkpoint_load: '%s' is truncated", path);
  int r = gfsb200_particles_upload (c, n, col[0].data (), col[1].data (), h.dim == 3 ? col[2].data () : NULL,
				    col[3].data (), col[4].data (), h.dim == 3 ? col[5].data () : NULL,
				    col[6].data (), col[7].data (), id.data ());
  if (r) return r;
  if (h.has_force && n) {
    if ((r = ensure_aux (c, n))) return r;
    for (int k = 0; k < 3; k++)
      CK (cudaMemcpyAsync (c->force[k], col[8 + k].data (), n*sizeof (double), cudaMemcpyHostToDevice, c->stream));
    CK (cudaStreamSynchronize (c->stream));
    c->forces_recorded = true;           /* a later sort or cull must carry the restored forces along */
  }
  return GFSB200_OK;
}

/* ------------------------------------------------------------------ */
/* two-way coupling                                                     */

/* The target of a deposit pass: the selected buffer, zeroed as the caller's mode requires.
 * Without a communicator: the components in `what' (bit 0 void fraction, bit 1 forces) are zeroed
 * here -- gfs_cell_reset on the leaves.  With one, comm.cu owns the policy (which buffer, which
 * slice is this rank's, what has to be waited for): gfsb200_comm_prepare_deposit. */
static int deposit_target (gfsb200_ctx * c, int what, bool local_only, DevDeposit * D)
{
  memset (D, 0, sizeof *D);
  const size_t n = c->T.n_cells;
  if (c->comm) {
    int r = gfsb200_comm_prepare_deposit (c->comm, what, local_only, D);
    if (r) return r;
  }
  else {
    if (what & 1) CK (cudaMemsetAsync (c->deposit, 0, n*sizeof (double), c->stream));
    if (what & 2) CK (cudaMemsetAsync (c->deposit + n, 0, (size_t) c->T.dim*n*sizeof (double), c->stream));
    D->local = c->deposit;
    D->own_lo = 0; D->own_hi = (int32_t) n;
    D->peers = NULL;
    c->dep_result = c->dep_which;
  }
  D->n_cells = (int64_t) n;
  return GFSB200_OK;
}

/* what: bit 0 = void fraction (component 0), bit 1 = force (components 1..dim) */
static int deposit (gfsb200_ctx * c, const gfsb200_step_params * p, int what)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "deposit: no tree resident");
  if ((what & 2) && !c->have_field) return gfsb200_fail (GFSB200_ERR_STATE, "deposit: no field resident");
  DevStep S;
  memset (&S, 0, sizeof S);
  if (what & 2) {
    int r = make_step (p, &S);
    if (r) return r;
  }
  CK (cudaSetDevice (c->device));
  if (what & 2) {
    int r = prepare_inertial (c, &S);
    if (r) return r;
  }
  DevDeposit D;
  int r = deposit_target (c, what, false, &D);
  if (r) return r;
  DevParticles P = particles_view (c);
  gfsb200_launch_deposit (&c->T, &c->F, &P, &S, what, &D, 0, NULL, NULL, c->stream);
  CK (cudaGetLastError ());
  return GFSB200_OK;
}

extern "C" int gfsb200_deposit_volume (gfsb200_ctx * c) { return deposit (c, NULL, 1); }

extern "C" int gfsb200_deposit_force (gfsb200_ctx * c, const gfsb200_step_params * p)
{
  return deposit (c, p, 2);
}

extern "C" int gfsb200_deposit_all (gfsb200_ctx * c, const gfsb200_step_params * p)
{
  return deposit (c, p, 3);
}

extern "C" int gfsb200_deposit_force_smoothed (gfsb200_ctx * c, const gfsb200_step_params * p,
						double rkernel, const gfsb200_kernel * kernel)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "deposit_force_smoothed: no tree resident");
  if (!c->have_field) return gfsb200_fail (GFSB200_ERR_STATE, "deposit_force_smoothed: no field resident");
  if (!kernel) return gfsb200_fail (GFSB200_ERR_ARG, "deposit_force_smoothed: null kernel");
  if (kernel->kind < GFSB200_KERNEL_CONSTANT || kernel->kind > GFSB200_KERNEL_COMPACT)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED,
			 "deposit_force_smoothed: kernel kind %d cannot be evaluated on the device", kernel->kind);
  if (kernel->kind == GFSB200_KERNEL_COMPACT && (kernel->p < 0 || kernel->p > 64))
    return gfsb200_fail (GFSB200_ERR_ARG, "deposit_force_smoothed: exponent %d out of range", kernel->p);
  if (!(rkernel >= 0.)) return gfsb200_fail (GFSB200_ERR_ARG, "deposit_force_smoothed: rkernel must be >= 0");
  /* the traversal keeps its path in 64 bits: GFSB200_MAX_LEVEL (20) levels x 3 bits fit */
  DevStep S;
  int r = make_step (p, &S);
  if (r) return r;
  CK (cudaSetDevice (c->device));
  if ((r = prepare_inertial (c, &S))) return r;
  if ((r = ensure_aux (c, c->n))) return r;
  c->knorm_n = 0;
  if (kernel->record_norm) {
    if (c->n > c->knorm_cap) {
      cudaFree (c->knorm); c->knorm = NULL; c->knorm_cap = 0;
      CK (cudaMalloc ((void **) &c->knorm, 2*(size_t) c->n*sizeof (double)));
      c->knorm_cap = c->n;
    }
    c->knorm_n = c->n;
  }
  const size_t n = c->T.n_cells;
  /* the kernel support of a particle reaches cells of any rank's slice: with a communicator this
     deposit is local and the exchange must be the all-reduce (GFSB200_EXCHANGE_ALLREDUCE) */
  DevDeposit D;
  if ((r = deposit_target (c, 2, true, &D))) return r;
  DevParticles P = particles_view (c);
  gfsb200_launch_deposit_smoothed (&c->T, &c->F, &P, &S, rkernel, kernel, D.local + n, D.local + 2*n,
				   c->T.dim == 3 ? D.local + 3*n : NULL,
				   kernel->record_norm ? c->knorm : NULL, c->stream);
  CK (cudaGetLastError ());
  c->forces_recorded = true;           /* the on-fluid forces (compute_forces_onfluid) are left in force[] */
  return GFSB200_OK;
}

extern "C" int gfsb200_download_kernel_norm (gfsb200_ctx * c, double * correction, double * volume)
{
  if (!c || !c->knorm_n) return gfsb200_fail (GFSB200_ERR_STATE, "download_kernel_norm: no recorded normalisation");
  CK (cudaSetDevice (c->device));
  if (correction)
    CK (cudaMemcpyAsync (correction, c->knorm, c->knorm_n*sizeof (double), cudaMemcpyDeviceToHost, c->stream));
  if (volume)
    CK (cudaMemcpyAsync (volume, c->knorm + c->knorm_n, c->knorm_n*sizeof (double), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return GFSB200_OK;
}

extern "C" int gfsb200_deposit_select (gfsb200_ctx * c, int which)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "deposit_select: no tree resident");
  if (which != 0 && which != 1) return gfsb200_fail (GFSB200_ERR_ARG, "deposit_select: buffer 0 or 1");
  CK (cudaSetDevice (c->device));
  if (!c->deposit_buf[which]) {
    CK (cudaMalloc ((void **) &c->deposit_buf[which], (size_t) c->deposit_count*sizeof (double)));
    CK (cudaMemsetAsync (c->deposit_buf[which], 0, (size_t) c->deposit_count*sizeof (double), c->stream));
  }
  c->deposit = c->deposit_buf[which];
  c->dep_which = which;
  if (!c->comm) c->dep_result = which;
  return GFSB200_OK;
}

extern "C" int gfsb200_deposit_buffer (gfsb200_ctx * c, double ** dev, int64_t * count)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "deposit_buffer: no tree resident");
  if (dev) *dev = c->comm ? c->deposit_buf[c->dep_result] : c->deposit;
  if (count) *count = c->deposit_count;
  return GFSB200_OK;
}

extern "C" int gfsb200_download_deposit (gfsb200_ctx * c, int comp, double * out)
{
  if (!c || !c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "download_deposit: no tree resident");
  if (comp < 0 || comp > c->T.dim || !out) return gfsb200_fail (GFSB200_ERR_ARG, "download_deposit: bad argument");
  CK (cudaSetDevice (c->device));
  const size_t n = c->T.n_cells;
  const double * src = c->deposit;
  if (c->comm) {
    /* the buffer whose exchange was issued last, once every rank's share has landed */
    int r = gfsb200_deposit_wait (c->comm);
    if (r) return r;
    src = c->deposit_buf[c->dep_result];
  }
  CK (cudaMemcpyAsync (out, src + comp*n, n*sizeof (double), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  if (c->comm) return gfsb200_comm_check (c->comm);
  return GFSB200_OK;
}

/* ------------------------------------------------------------------ */

extern "C" int gfsb200_timer_reset (gfsb200_ctx * c)
{
  if (!c) return gfsb200_fail (GFSB200_ERR_ARG, "null context");
  CK (cudaSetDevice (c->device));
  CK (cudaStreamSynchronize (c->stream));
  c->ev_used = 0;
  c->timer_calls = 0;          /* the first launch after a reset is a timed one */
  return GFSB200_OK;
}

extern "C" int gfsb200_timer_sampling (gfsb200_ctx * c, int every)
{
  if (!c) return gfsb200_fail (GFSB200_ERR_ARG, "null context");
  c->timer_every = every;
  c->timer_calls = 0;
  return GFSB200_OK;
}

extern "C" int gfsb200_timer_read (gfsb200_ctx * c, double * step_kernel_ms, int64_t * launches)
{
  if (!c) return gfsb200_fail (GFSB200_ERR_ARG, "null context");
  CK (cudaSetDevice (c->device));
  CK (cudaStreamSynchronize (c->stream));
  double total = 0.;
  for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
    float ms = 0.f;
    CK (cudaEventElapsedTime (&ms, c->ev[i], c->ev[i + 1]));
    total += ms;
  }
  const int64_t n = (int64_t) (c->ev_used/2);
  if (step_kernel_ms) *step_kernel_ms = n ? total/n : 0.;
  if (launches) *launches = n;
  return GFSB200_OK;
}
