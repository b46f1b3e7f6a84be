/* ctx_internal.cuh -- the device context shared by capi.cu and comm.cu (not part of the C-ABI) */
#ifndef GFSB200_CTX_INTERNAL_CUH
#define GFSB200_CTX_INTERNAL_CUH

#include <cuda_runtime.h>
#include <vector>
#include "gfsb200_internal.h"
#include "device_types.cuh"

struct gfsb200_comm;

#define NCOL 8     /* x y z vx vy vz mass volume */

struct gfsb200_ctx {
  int device, n_sm;
  cudaStream_t stream;
  cudaStream_t aux_stream;     /* second stream of the cell pass */
  cudaEvent_t ev_fork, ev_join;
  /* tree */
  bool have_tree;
  DevTree T;
  int32_t * d_child0, * d_neighbor, * d_la_slot, * d_vtx_off, * d_vtx_cell, * d_leaf_vtx, * d_parent;
  double * d_solid_a, * d_solid_s;
  uint8_t * d_level, * d_info;
  double * d_vtx_w, * d_vtx_wuni;
  /* field */
  bool have_field, own_field;
  DevField F;
  double * d_field[5];         /* owned copies: u v w alpha mu */
  double * d_prev[3];          /* Un Vn Wn (GfsForceInertial / GfsForceAddedMass) */
  bool have_prev, acc_valid;
  /* particles */
  int64_t n, cap;
  double * col[2][NCOL];       /* double-buffered SoA */
  uint32_t * id[2];
  int cur;
  double * force[3];
  int32_t * cell;
  int64_t aux_cap;             /* capacity of force/cell/perm/key buffers */
  int32_t * perm, * perm2;
  uint32_t * key, * key2;
  uint8_t * flag;
  int32_t * d_count;
  void * cub_tmp;
  size_t cub_tmp_bytes;
  double ** d_ptr_table;       /* [2][NCOL] device copy of col pointers */
  /* escape tracking for gfs_particle_bc */
  bool mark_outside;           /* the next tracked step flags (c->flag) the particles outside before it */
  bool forces_recorded;        /* force[] holds what the last step (or on-fluid pass) recorded for the resident list */
  int * esc_count;             /* [4]: escaped, wrapped, dropped, outside the domain before the step */
  int32_t * esc_idx;
  double * esc_old;
  int esc_cap;
  bool esc_armed;              /* the last step tracked escapes */
  bool last_step_fused;        /* the last step also deposited (fuse_deposit) */
  DevStep last_S;              /* its parameters and target, for the deposit of the particles gfs_particle_bc wraps */
  DevDeposit last_D;
  /* host-list pipeline (gfsb200_step_host) */
  double * hp_col[3][NCOL];
  int64_t hp_chunk;
  cudaStream_t hp_h2d, hp_d2h;
  cudaEvent_t hp_in[3], hp_done[3], hp_out[3];
  /* deposit: two buffers so that the all-reduce of one step can overlap the next step */
  double * deposit;            /* the selected one: target of the deposit calls */
  double * deposit_buf[2];
  int dep_which;               /* index of the selected buffer */
  int dep_result;              /* buffer the download reads: the selected one, or with a communicator the
				  one whose exchange was issued last */
  int64_t deposit_count;
  gfsb200_comm * comm;         /* attached by gfsb200_comm_init_*; owns the exchange of the deposit */
  int64_t tree_generation;     /* bumped by gfsb200_upload_tree (the deposit buffers are reallocated) */
  int32_t leaf_lo, leaf_hi;    /* the leaves of the GfsBox trees -- the only cells a deposit can reach -- lie in [leaf_lo, leaf_hi) */
  double * scratch;            /* pooled device scratch of the batched point queries */
  size_t scratch_bytes;
  double * knorm;              /* [2][knorm_n] correction, volume of the last smoothed deposit */
  int64_t knorm_cap, knorm_n;
  int step_minb;               /* __launch_bounds__ min blocks/SM variant of the step kernel */
  int step_mode;               /* gfsb200_launch_step's mode; < 0: chosen by tree type */
  /* timing */
  std::vector<cudaEvent_t> ev;
  size_t ev_used;
  bool timing;
  int timer_every;             /* gfsb200_timer_sampling: events around every n-th launch (1: all, 0: none) */
  long timer_calls;
};

/* comm.cu: the exchange of the deposited field between the GPUs of one box */
extern "C" {
int gfsb200_comm_prepare_deposit (gfsb200_comm * m, int what, bool local_only, DevDeposit * D);
void gfsb200_comm_tree_changed (gfsb200_comm * m);
void gfsb200_comm_detach (gfsb200_comm * m);
int gfsb200_comm_check (gfsb200_comm * m);
/* capi.cu internals comm.cu builds on */
int gfsb200_internal_sort (gfsb200_ctx * c);                 /* sort by cell; sorted keys left in c->key2 */
int gfsb200_internal_sort_by_owner (gfsb200_ctx * c, const uint8_t * owner_of, int nranks);
int gfsb200_internal_reserve (gfsb200_ctx * c, int64_t n);   /* capacity for n particles, contents kept */
int gfsb200_internal_ensure_aux (gfsb200_ctx * c, int64_t n);
int gfsb200_internal_field_buffers (gfsb200_ctx * c, const int present[5]);
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return gfsb200_fail (GFSB200_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString (e_), __FILE__, __LINE__); \
  } while (0)


#endif
