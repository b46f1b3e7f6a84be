/* comm.cu -- the particulate path over the GPUs of one box.
 *
 * Particles shard over the GPUs; the flat tree and the field are replicated (north_star).  One-way
 * coupling needs no exchange at all.  Two-way coupling has exactly one: every rank must end the
 * step with sum_r deposit_r, the field GfsParticulateField / GfsSourceParticulate hand to the fluid
 * solver.  This file owns
 *
 *   gfsb200_comm_rebalance     the device-side replacement of the reference's per-particle MPI
 *                              migration (mpi_send_particle / mpi_rcv_particle,
 *                              modules/particulatecommon.c:3218-3244; gfs_send_objects,
 *                              src/domain.c:4464-4557): one histogram all-reduce fixes slice
 *                              boundaries in the global cell order, one grouped send/recv moves
 *                              every particle to the rank that owns its cell;
 *   gfsb200_deposit_allreduce  the sum of the deposited field (the reference reduces scalars only,
 *                              gfs_all_reduce, src/utils.h:36-42).  After a rebalance the deposit
 *                              kernels reduce every contribution straight into the OWNER's slice
 *                              (local L2 atomics; remote fp64 reductions over NVLink peer memory
 *                              for the drifters -- run_deposit / owner_base in
 *                              particle_kernels.cu), so the exchange is a cross-GPU barrier (flags
 *                              in peer memory), a push of the own slice to every peer by the copy
 *                              engines and a completion flag: an all-gather, (R-1)/R of the field
 *                              per GPU, no reduction pass and no SM time.  Otherwise (no peer
 *                              access, no rebalance, smoothed deposit): ncclAllReduce;
 *   gfsb200_broadcast_field    U,V,W over PCIe once, then NVLink.
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2): libgfsb200.so has no link-time dependency on
 * it and a process that already carries an NCCL (torch) shares that copy.
 *
 * Ordering of the owner-slice protocol (two deposit buffers, step n uses buffer n % 2):
 *   D(n)    deposit: local + remote reductions into n%2 (the context's stream)
 *   Z(n+1)  zero the OWN slice of the other buffer: first thing on the communication stream after D(n)
 *   X(n)    on the communication stream, after D(n): barrier B(n) -- every rank has finished D(n),
 *           so every remote reduction into my slice has landed (kernel completion drains them) --
 *           then the pushes of my slice, then the DONE flags.
 * The context's stream orders D(n+1) after B(n) (an event recorded right behind the barrier kernel:
 * the pushes are not waited for).  So a peer's D(n+1) reduces into my slice of buffer (n+1)%2 only
 * after it has passed B(n), which waits for my arrival, which follows my Z(n+1): the slice is zero
 * before anything lands in it.  My Z(n+2) waits for my X(n): the slice is not cleared while it is
 * being pushed.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>              /* types and prototypes only: the library is bound with dlsym */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <vector>
#include "ctx_internal.cuh"

extern "C" long long gfsb200_launch_counter;

/* ------------------------------------------------------------------ */
/* NCCL, bound at run time                                              */

namespace {

struct Nccl {
  void * handle;
  decltype (&ncclGetUniqueId) GetUniqueId;
  decltype (&ncclCommInitRank) CommInitRank;
  decltype (&ncclCommInitAll) CommInitAll;
  decltype (&ncclCommDestroy) CommDestroy;
  decltype (&ncclAllReduce) AllReduce;
  decltype (&ncclBroadcast) Broadcast;
  decltype (&ncclAllGather) AllGather;
  decltype (&ncclSend) Send;
  decltype (&ncclRecv) Recv;
  decltype (&ncclGroupStart) GroupStart;
  decltype (&ncclGroupEnd) GroupEnd;
  decltype (&ncclGetErrorString) GetErrorString;
};

Nccl * nccl_api ()
{
  static Nccl N;
  static int state = 0;        /* 0 not tried, 1 bound, -1 unavailable */
  if (state == 0) {
    const char * names[] = { getenv ("GFSB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    N.handle = NULL;
    for (int k = 0; k < 3 && !N.handle; k++)
      if (names[k]) N.handle = dlopen (names[k], RTLD_NOW | RTLD_GLOBAL);
    state = -1;
    if (N.handle) {
      bool ok = true;
#define BIND(f) do { N.f = (decltype (N.f)) dlsym (N.handle, "nccl" #f); if (!N.f) ok = false; } while (0)
      BIND (GetUniqueId); BIND (CommInitRank); BIND (CommInitAll); BIND (CommDestroy); BIND (AllReduce);
      BIND (Broadcast); BIND (AllGather); BIND (Send); BIND (Recv); BIND (GroupStart); BIND (GroupEnd);
      BIND (GetErrorString);
#undef BIND
      if (ok) state = 1;
    }
  }
  return state == 1 ? &N : NULL;
}

#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
  return gfsb200_fail (GFSB200_ERR_CUDA, "%s: %s (%s:%d)", #call, nccl_api ()->GetErrorString (r_), __FILE__, __LINE__); \
  } while (0)

/* ------------------------------------------------------------------ */
/* flags in peer memory                                                 */

enum { FLAG_ARRIVE = 0, FLAG_DONE = 1, FLAG_ERR = 2, FLAG_SRC = 3, FLAG_ROWS = 4 };
#define FLAGS_BYTES (2u << 20)       /* its own allocation block (cudaIpc maps whole blocks) */
#define SPIN_TIMEOUT_NS 20000000000ull

struct PeerFlags { uint32_t * p[GFSB200_MAX_RANKS]; };

__device__ __forceinline__ unsigned long long global_ns ()
{
  unsigned long long t;
  asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void st_release_sys (uint32_t * p, uint32_t v)
{
  asm volatile ("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_sys (const uint32_t * p)
{
  uint32_t v;
  asm volatile ("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

/* thread t: (signal) write `epoch' into row `row', column `self' of rank t's flags; (wait) spin
 * until rank t has written an epoch >= `epoch' into the same row of MY flags.  One block. */
__global__ void comm_flags_kernel (PeerFlags peers, uint32_t * mine, int self, int n, int row,
				   uint32_t epoch, int signal, int wait)
{
  const int t = threadIdx.x;
  if (t >= n) return;
  if (t == 0)                  /* the value the copy engine hands to the peers as DONE flag, later in the stream */
    mine[FLAG_SRC*GFSB200_MAX_RANKS] = epoch;
  if (signal) {
    __threadfence_system ();
    st_release_sys (peers.p[t] + row*GFSB200_MAX_RANKS + self, epoch);
  }
  if (wait) {
    const unsigned long long t0 = global_ns ();
    while ((int32_t) (ld_acquire_sys (mine + row*GFSB200_MAX_RANKS + t) - epoch) < 0) {
      if (global_ns () - t0 > SPIN_TIMEOUT_NS) {       /* a peer never arrived: report, do not hang the GPU */
	mine[FLAG_ERR*GFSB200_MAX_RANKS] = 1u + (uint32_t) t;
	break;
      }
      __nanosleep (200);
    }
  }
}

/* ------------------------------------------------------------------ */
/* rebalance kernels                                                    */

__global__ void __launch_bounds__(256)
cell_histogram_kernel (int64_t n, const uint32_t * __restrict__ sorted_key, uint32_t * __restrict__ hist)
{
  /* keys are sorted: one atomic per run of equal keys inside a warp */
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const uint32_t key = i < n ? sorted_key[i] : 0xffffffffu;
  const uint32_t prev = __shfl_up_sync (full, key, 1);
  const bool head = lane == 0 || prev != key;
  const unsigned heads = __ballot_sync (full, head);
  const unsigned above = lane == 31 ? 0u : heads & (0xffffffffu << (lane + 1));
  const int run_end = above ? __ffs (above) - 1 : 32;
  if (head && i < n)
    atomicAdd (hist + key, (uint32_t) (run_end - lane));
}

/* pos[k] = first index whose sorted key is >= bound[k] */
__global__ void lower_bound_kernel (int64_t n, const uint32_t * __restrict__ sorted_key, int nb,
				    const uint32_t * __restrict__ bound, int32_t * __restrict__ pos)
{
  const int k = threadIdx.x;
  if (k >= nb) return;
  int64_t lo = 0, hi = n;
  const uint32_t b = bound[k];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (sorted_key[mid] < b) lo = mid + 1; else hi = mid;
  }
  pos[k] = (int32_t) lo;
}

} // namespace

/* ------------------------------------------------------------------ */

#define N_PUSH 4

struct PeerInfo {
  uint64_t host;
  int32_t pid, device, rank, pad;
  uint64_t ptr[3], off[3];     /* deposit buffer 0, deposit buffer 1, flags: address and offset in its block */
  cudaIpcMemHandle_t handle[3];
};

struct gfsb200_comm {
  gfsb200_ctx * c;
  int rank, nranks;
  ncclComm_t nccl;
  int exchange_mode;           /* GFSB200_EXCHANGE_* */
  cudaStream_t stream;         /* the communication stream */
  cudaStream_t push[N_PUSH];   /* the pushes of one exchange fan out over these */
  cudaEvent_t ev_fork, ev_join[N_PUSH];
  cudaEvent_t ev_dep, ev_x[2];
  cudaEvent_t ev_barrier;      /* B(n) passed: every rank has cleared its slices for step n+1 */
  bool barrier_pending;
  bool x_pending[2];           /* an exchange of buffer b has been issued and not yet waited for by the context's stream */
  int x_mode[2];
  uint32_t x_epoch[2];
  int dep_mode;                /* mode of the deposits into the target since its last exchange: 0 none, 1 owner, 2 local */
  int dep_done;                /* components deposited since then (bit 0 void fraction, bit 1 forces) */
  bool ahead_ok[2];            /* the own slice of buffer b is zero and nothing has been deposited into it */
  /* peers */
  bool p2p;                    /* every rank reaches every other rank's memory */
  uint32_t * flags;
  uint32_t * peer_flags[GFSB200_MAX_RANKS];
  double * peer_dep[2][GFSB200_MAX_RANKS];
  void * ipc_mapped[3*GFSB200_MAX_RANKS];
  int n_mapped;
  int64_t peer_generation;
  DevOwners * d_owners;        /* [2]: one table per deposit buffer */
  /* ownership */
  bool owner_valid;
  int32_t split[GFSB200_MAX_RANKS + 1];
  double share_cum[GFSB200_MAX_RANKS + 1];   /* gfsb200_comm_set_shares: cumulative target shares (share_cum[R] = 1) */
  bool shares_set;
  bool by_owner;               /* adaptive trees: the slices are ranges of the depth-first leaf order (owner table) */
  uint8_t * d_owner_of; int64_t owner_cap;
  std::vector<uint8_t> owner_of;             /* host copy */
  std::vector<int32_t> range_lo, range_hi;   /* the cell ranges that hold this rank's leaves (one per level at most) */
  uint32_t epoch;
  /* scratch */
  uint32_t * d_hist; int64_t hist_cap;
  int32_t * d_small;           /* bounds, positions, count matrix */
  PeerInfo * d_info;
  /* statistics */
  std::vector<cudaEvent_t> tev;
  size_t tev_used;
  int64_t bytes_sent;
};

#define MODE_OWNER 1
#define MODE_LOCAL 2
#define SMALL_INTS (4*GFSB200_MAX_RANKS + GFSB200_MAX_RANKS*GFSB200_MAX_RANKS + 16)

static uint64_t host_id ()
{
  char name[256] = { 0 };
  gethostname (name, sizeof name - 1);
  uint64_t h = 1469598103934665603ull;
  for (const char * p = name; *p; p++) { h ^= (unsigned char) *p; h *= 1099511628211ull; }
  FILE * f = fopen ("/proc/sys/kernel/random/boot_id", "r");
  if (f) {
    int ch;
    while ((ch = fgetc (f)) != EOF) { h ^= (unsigned char) ch; h *= 1099511628211ull; }
    fclose (f);
  }
  return h;
}

/* base address of the cudaMalloc block that holds p (cuMemGetAddressRange, bound through the
 * runtime so that the library needs no libcuda at link time) */
static int block_base (const void * p, uint64_t * base)
{
  typedef int (* range_fn) (unsigned long long *, size_t *, unsigned long long);
  static range_fn fn = NULL;
  if (!fn) {
    void * sym = NULL;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint ("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym)
      return gfsb200_fail (GFSB200_ERR_CUDA, "cuMemGetAddressRange is not available");
    fn = (range_fn) sym;
  }
  unsigned long long b = 0; size_t sz = 0;
  if (fn (&b, &sz, (unsigned long long) (uintptr_t) p) != 0)
    return gfsb200_fail (GFSB200_ERR_CUDA, "cuMemGetAddressRange failed");
  *base = b;
  return GFSB200_OK;
}

static void close_peers (gfsb200_comm * m)
{
  for (int k = 0; k < m->n_mapped; k++)
    cudaIpcCloseMemHandle (m->ipc_mapped[k]);
  m->n_mapped = 0;
  m->peer_generation = -1;
}

static int sync_all (gfsb200_comm * const * local, int n_local)
{
  for (int k = 0; k < n_local; k++) {
    CK (cudaSetDevice (local[k]->c->device));
    CK (cudaStreamSynchronize (local[k]->stream));
    CK (cudaStreamSynchronize (local[k]->c->stream));
  }
  return GFSB200_OK;
}

static int check_local (gfsb200_comm * const * local, int n_local, const char * what)
{
  if (!local || n_local <= 0)
    return gfsb200_fail (GFSB200_ERR_ARG, "%s: no communicator", what);
  for (int k = 0; k < n_local; k++)
    if (!local[k] || !local[k]->c)
      return gfsb200_fail (GFSB200_ERR_ARG, "%s: communicator %d is null or detached", what, k);
  return GFSB200_OK;
}

/* Collective: every rank learns where every other rank keeps its deposit buffers and flags, and
 * maps them (same process: peer access; another process of the same box: cudaIpc). */
static int exchange_peers (gfsb200_comm * const * local, int n_local)
{
  Nccl * N = nccl_api ();
  const int R = local[0]->nranks;
  std::vector<PeerInfo> all ((size_t) n_local*R);
  const uint64_t host = host_id ();
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    gfsb200_ctx * c = m->c;
    CK (cudaSetDevice (c->device));
    close_peers (m);
    if (c->have_tree && !c->deposit_buf[1]) {
      CK (cudaMalloc ((void **) &c->deposit_buf[1], (size_t) c->deposit_count*sizeof (double)));
      CK (cudaMemsetAsync (c->deposit_buf[1], 0, (size_t) c->deposit_count*sizeof (double), c->stream));
    }
    PeerInfo me;
    memset (&me, 0, sizeof me);
    me.host = host; me.pid = (int32_t) getpid (); me.device = c->device; me.rank = m->rank;
    void * ptr[3] = { c->deposit_buf[0], c->deposit_buf[1], m->flags };
    for (int j = 0; j < 3; j++) {
      me.ptr[j] = (uint64_t) (uintptr_t) ptr[j];
      if (ptr[j] && R > 1) {
	uint64_t base = 0;
	int r = block_base (ptr[j], &base);
	if (r) return r;
	me.off[j] = me.ptr[j] - base;
	CK (cudaIpcGetMemHandle (&me.handle[j], ptr[j]));
      }
    }
    CK (cudaMemcpyAsync (m->d_info + m->rank, &me, sizeof me, cudaMemcpyHostToDevice, c->stream));
  }
  if (R > 1) {
    NK (N->GroupStart ());
    for (int k = 0; k < n_local; k++) {
      gfsb200_comm * m = local[k];
      NK (N->AllGather (m->d_info + m->rank, m->d_info, sizeof (PeerInfo), ncclChar, m->nccl, m->c->stream));
    }
    NK (N->GroupEnd ());
  }
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    CK (cudaSetDevice (m->c->device));
    CK (cudaMemcpyAsync (&all[(size_t) k*R], m->d_info, R*sizeof (PeerInfo), cudaMemcpyDeviceToHost, m->c->stream));
    CK (cudaStreamSynchronize (m->c->stream));
  }
  int ok_all = 1;
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    gfsb200_ctx * c = m->c;
    const PeerInfo * info = &all[(size_t) k*R];
    CK (cudaSetDevice (c->device));
    bool ok = true;
    for (int q = 0; q < R; q++) {
      void * mapped[3] = { NULL, NULL, NULL };
      if (q == m->rank) {
	mapped[0] = c->deposit_buf[0]; mapped[1] = c->deposit_buf[1]; mapped[2] = m->flags;
      }
      else if (info[q].host != host)
	ok = false;                       /* another box: NCCL only */
      else if (info[q].pid == (int32_t) getpid ()) {
	int can = 0;
	if (cudaDeviceCanAccessPeer (&can, c->device, info[q].device) != cudaSuccess || !can)
	  ok = false;
	else {
	  cudaError_t e = cudaDeviceEnablePeerAccess (info[q].device, 0);
	  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
	  cudaGetLastError ();
	  for (int j = 0; j < 3; j++) mapped[j] = (void *) (uintptr_t) info[q].ptr[j];
	}
      }
      else {
	for (int j = 0; j < 3 && ok; j++) {
	  if (!info[q].ptr[j]) continue;
	  /* two pointers of one peer may live in the same block: map it once */
	  void * base = NULL;
	  for (int i = 0; i < j; i++)
	    if (info[q].ptr[i] && !memcmp (&info[q].handle[i], &info[q].handle[j], sizeof (cudaIpcMemHandle_t)))
	      base = (char *) mapped[i] - info[q].off[i];
	  if (!base) {
	    if (cudaIpcOpenMemHandle (&base, info[q].handle[j], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
	      cudaGetLastError ();
	      ok = false;
	      break;
	    }
	    m->ipc_mapped[m->n_mapped++] = base;
	  }
	  mapped[j] = (char *) base + info[q].off[j];
	}
      }
      m->peer_dep[0][q] = (double *) mapped[0];
      m->peer_dep[1][q] = (double *) mapped[1];
      m->peer_flags[q] = (uint32_t *) mapped[2];
    }
    m->p2p = ok;
    if (!ok) ok_all = 0;
    m->peer_generation = c->tree_generation;
  }
  /* every rank must take the same path: peer access counts only if all of them have it */
  if (R > 1) {
    for (int k = 0; k < n_local; k++) {
      gfsb200_comm * m = local[k];
      CK (cudaSetDevice (m->c->device));
      int32_t v = ok_all && m->p2p ? 1 : 0;
      CK (cudaMemcpyAsync (m->d_small, &v, sizeof v, cudaMemcpyHostToDevice, m->c->stream));
    }
    NK (N->GroupStart ());
    for (int k = 0; k < n_local; k++)
      NK (N->AllReduce (local[k]->d_small, local[k]->d_small, 1, ncclInt32, ncclMin, local[k]->nccl, local[k]->c->stream));
    NK (N->GroupEnd ());
    for (int k = 0; k < n_local; k++) {
      gfsb200_comm * m = local[k];
      int32_t v = 0;
      CK (cudaSetDevice (m->c->device));
      CK (cudaMemcpyAsync (&v, m->d_small, sizeof v, cudaMemcpyDeviceToHost, m->c->stream));
      CK (cudaStreamSynchronize (m->c->stream));
      m->p2p = v != 0;
    }
  }
  if (getenv ("GFSB200_NO_P2P"))
    for (int k = 0; k < n_local; k++) local[k]->p2p = false;
  return GFSB200_OK;
}

static int comm_new (gfsb200_ctx * c, int rank, int nranks, ncclComm_t nc, gfsb200_comm ** out)
{
  gfsb200_comm * m = new gfsb200_comm ();
  m->c = c; m->rank = rank; m->nranks = nranks; m->nccl = nc;
  m->exchange_mode = getenv ("GFSB200_EXCHANGE") ? atoi (getenv ("GFSB200_EXCHANGE")) : GFSB200_EXCHANGE_AUTO;
  m->x_pending[0] = m->x_pending[1] = false;
  m->x_mode[0] = m->x_mode[1] = 0;
  m->x_epoch[0] = m->x_epoch[1] = 0;
  m->dep_mode = m->dep_done = 0;
  m->ahead_ok[0] = m->ahead_ok[1] = false;
  m->p2p = nranks == 1;
  m->n_mapped = 0; m->peer_generation = -1;
  m->owner_valid = false;
  m->epoch = 0;
  m->d_hist = NULL; m->hist_cap = 0;
  m->by_owner = false; m->d_owner_of = NULL; m->owner_cap = 0;
  m->shares_set = false;
  m->tev_used = 0; m->bytes_sent = 0;
  memset (m->peer_flags, 0, sizeof m->peer_flags);
  memset (m->peer_dep, 0, sizeof m->peer_dep);
  CK (cudaSetDevice (c->device));
  int lo = 0, hi = 0;
  CK (cudaDeviceGetStreamPriorityRange (&lo, &hi));
  CK (cudaStreamCreateWithPriority (&m->stream, cudaStreamNonBlocking, hi));
  for (int j = 0; j < N_PUSH; j++) {
    CK (cudaStreamCreateWithPriority (&m->push[j], cudaStreamNonBlocking, hi));
    CK (cudaEventCreateWithFlags (&m->ev_join[j], cudaEventDisableTiming));
  }
  CK (cudaEventCreateWithFlags (&m->ev_fork, cudaEventDisableTiming));
  CK (cudaEventCreateWithFlags (&m->ev_dep, cudaEventDisableTiming));
  CK (cudaEventCreateWithFlags (&m->ev_barrier, cudaEventDisableTiming));
  m->barrier_pending = false;
  CK (cudaEventCreateWithFlags (&m->ev_x[0], cudaEventDisableTiming));
  CK (cudaEventCreateWithFlags (&m->ev_x[1], cudaEventDisableTiming));
  CK (cudaMalloc ((void **) &m->flags, FLAGS_BYTES));
  CK (cudaMemset (m->flags, 0, FLAGS_BYTES));
  CK (cudaMalloc ((void **) &m->d_owners, 2*sizeof (DevOwners)));
  CK (cudaMemset (m->d_owners, 0, 2*sizeof (DevOwners)));
  CK (cudaMalloc ((void **) &m->d_small, SMALL_INTS*sizeof (int32_t)));
  CK (cudaMalloc ((void **) &m->d_info, GFSB200_MAX_RANKS*sizeof (PeerInfo)));
  m->peer_flags[rank] = m->flags;
  c->comm = m;
  *out = m;
  return GFSB200_OK;
}

extern "C" int gfsb200_comm_unique_id (void * id)
{
  if (!id) return gfsb200_fail (GFSB200_ERR_ARG, "comm_unique_id: null argument");
  Nccl * N = nccl_api ();
  if (!N) return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "libnccl.so.2 cannot be loaded: %s", dlerror ());
  ncclUniqueId u;
  NK (N->GetUniqueId (&u));
  memcpy (id, &u, GFSB200_UNIQUE_ID_BYTES);
  return GFSB200_OK;
}

extern "C" int gfsb200_comm_init_rank (gfsb200_ctx * c, const void * id, int rank, int nranks,
				       gfsb200_comm ** out)
{
  if (!c || !out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id))
    return gfsb200_fail (GFSB200_ERR_ARG, "comm_init_rank: bad argument");
  if (nranks > GFSB200_MAX_RANKS)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "comm_init_rank: at most %d ranks (one box)", GFSB200_MAX_RANKS);
  if (c->comm) return gfsb200_fail (GFSB200_ERR_STATE, "comm_init_rank: the context already has a communicator");
  *out = NULL;
  CK (cudaSetDevice (c->device));
  ncclComm_t nc = NULL;
  if (nranks > 1) {
    Nccl * N = nccl_api ();
    if (!N) return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "libnccl.so.2 cannot be loaded: %s", dlerror ());
    ncclUniqueId u;
    memcpy (&u, id, GFSB200_UNIQUE_ID_BYTES);
    NK (N->CommInitRank (&nc, nranks, u, rank));
  }
  int r = comm_new (c, rank, nranks, nc, out);
  if (r) return r;
  return exchange_peers (out, 1);
}

extern "C" int gfsb200_comm_init_all (int n, gfsb200_ctx * const * ctxs, gfsb200_comm ** out)
{
  if (n < 1 || !ctxs || !out) return gfsb200_fail (GFSB200_ERR_ARG, "comm_init_all: bad argument");
  if (n > GFSB200_MAX_RANKS)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "comm_init_all: at most %d ranks (one box)", GFSB200_MAX_RANKS);
  std::vector<int> devs (n);
  for (int k = 0; k < n; k++) {
    if (!ctxs[k]) return gfsb200_fail (GFSB200_ERR_ARG, "comm_init_all: null context %d", k);
    if (ctxs[k]->comm) return gfsb200_fail (GFSB200_ERR_STATE, "comm_init_all: context %d already has a communicator", k);
    devs[k] = ctxs[k]->device;
    out[k] = NULL;
  }
  std::vector<ncclComm_t> nc (n, (ncclComm_t) NULL);
  if (n > 1) {
    Nccl * N = nccl_api ();
    if (!N) return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "libnccl.so.2 cannot be loaded: %s", dlerror ());
    NK (N->CommInitAll (nc.data (), n, devs.data ()));
  }
  for (int k = 0; k < n; k++) {
    int r = comm_new (ctxs[k], k, n, nc[k], &out[k]);
    if (r) return r;
  }
  return exchange_peers (out, n);
}

extern "C" void gfsb200_comm_detach (gfsb200_comm * m)
{
  if (!m || !m->c) return;
  cudaSetDevice (m->c->device);
  cudaStreamSynchronize (m->stream);
  close_peers (m);
  m->c->comm = NULL;
  m->c = NULL;
}

extern "C" void gfsb200_comm_destroy (gfsb200_comm * m)
{
  if (!m) return;
  int dev = -1;
  if (m->c) { dev = m->c->device; gfsb200_comm_detach (m); }
  if (dev >= 0) cudaSetDevice (dev);
  if (m->nccl && nccl_api ()) nccl_api ()->CommDestroy (m->nccl);
  cudaStreamDestroy (m->stream);
  for (int j = 0; j < N_PUSH; j++) { cudaStreamDestroy (m->push[j]); cudaEventDestroy (m->ev_join[j]); }
  cudaEventDestroy (m->ev_fork);
  cudaEventDestroy (m->ev_dep); cudaEventDestroy (m->ev_barrier); cudaEventDestroy (m->ev_x[0]); cudaEventDestroy (m->ev_x[1]);
  for (size_t i = 0; i < m->tev.size (); i++) cudaEventDestroy (m->tev[i]);
  cudaFree (m->flags); cudaFree (m->d_owners); cudaFree (m->d_small); cudaFree (m->d_info); cudaFree (m->d_hist);
  cudaFree (m->d_owner_of);
  delete m;
}

extern "C" int gfsb200_comm_rank (const gfsb200_comm * m) { return m ? m->rank : -1; }
extern "C" int gfsb200_comm_size (const gfsb200_comm * m) { return m ? m->nranks : -1; }
extern "C" int gfsb200_comm_peer_access (const gfsb200_comm * m) { return m && m->p2p ? 1 : 0; }

extern "C" int gfsb200_comm_set_exchange (gfsb200_comm * m, int mode)
{
  if (!m || (mode != GFSB200_EXCHANGE_AUTO && mode != GFSB200_EXCHANGE_ALLREDUCE))
    return gfsb200_fail (GFSB200_ERR_ARG, "comm_set_exchange: bad argument");
  if (m->dep_mode)
    return gfsb200_fail (GFSB200_ERR_STATE, "comm_set_exchange: a deposit is waiting for its gfsb200_deposit_allreduce");
  m->exchange_mode = mode;
  return GFSB200_OK;
}

/* the tree is about to be replaced: the deposit buffers are reallocated, slices and peer mappings lapse */
extern "C" void gfsb200_comm_tree_changed (gfsb200_comm * m)
{
  if (!m || !m->c) return;
  cudaStreamSynchronize (m->stream);
  close_peers (m);
  m->owner_valid = false;
  m->by_owner = false;
  m->x_pending[0] = m->x_pending[1] = false;
  m->barrier_pending = false;
  m->dep_mode = m->dep_done = 0;
  m->ahead_ok[0] = m->ahead_ok[1] = false;
}

extern "C" int gfsb200_comm_check (gfsb200_comm * m)
{
  if (!m || !m->c) return gfsb200_fail (GFSB200_ERR_ARG, "comm_check: null communicator");
  uint32_t err = 0;
  CK (cudaMemcpy (&err, m->flags + FLAG_ERR*GFSB200_MAX_RANKS, sizeof err, cudaMemcpyDeviceToHost));
  if (err)
    return gfsb200_fail (GFSB200_ERR_STATE, "rank %d waited more than %.0f s for rank %u in the deposit exchange",
			 m->rank, SPIN_TIMEOUT_NS*1e-9, err - 1);
  return GFSB200_OK;
}

/* ------------------------------------------------------------------ */
/* field                                                                */

extern "C" int gfsb200_internal_field_buffers (gfsb200_ctx * c, const int present[5]);

extern "C" int gfsb200_broadcast_field (gfsb200_comm * const * local, int n_local, int root,
					const double * u, const double * v, const double * w,
					const double * alpha, const double * mu)
{
  int r = check_local (local, n_local, "broadcast_field");
  if (r) return r;
  const int R = local[0]->nranks;
  if (root < 0 || root >= R) return gfsb200_fail (GFSB200_ERR_ARG, "broadcast_field: root %d out of range", root);
  Nccl * N = R > 1 ? nccl_api () : NULL;
  const double * src[5] = { u, v, w, alpha, mu };
  for (int k = 0; k < n_local; k++) {
    gfsb200_ctx * c = local[k]->c;
    if (!c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "broadcast_field: upload a tree first");
    CK (cudaSetDevice (c->device));
    int present[5] = { 1, 1, c->T.dim == 3, alpha != NULL, mu != NULL };
    if (local[k]->rank == root && (!u || !v || (c->T.dim == 3 && !w)))
      return gfsb200_fail (GFSB200_ERR_ARG, "broadcast_field: missing velocity component on the root");
    if ((r = gfsb200_internal_field_buffers (c, present))) return r;
    if (local[k]->rank == root)
      for (int i = 0; i < 5; i++)
	if (present[i])
	  CK (cudaMemcpyAsync (c->d_field[i], src[i], (size_t) c->T.n_cells*sizeof (double), cudaMemcpyHostToDevice, c->stream));
  }
  if (R > 1) {
    NK (N->GroupStart ());
    for (int k = 0; k < n_local; k++) {
      gfsb200_ctx * c = local[k]->c;
      for (int i = 0; i < 5; i++)
	if (c->d_field[i])
	  NK (N->Broadcast (c->d_field[i], c->d_field[i], (size_t) c->T.n_cells, ncclDouble, root, local[k]->nccl, c->stream));
    }
    NK (N->GroupEnd ());
  }
  for (int k = 0; k < n_local; k++)
    if ((r = gfsb200_refresh_field (local[k]->c))) return r;
  return GFSB200_OK;
}

/* ------------------------------------------------------------------ */
/* rebalance                                                            */

extern "C" int gfsb200_comm_splitters (const uint32_t * count, int32_t n_cells, int nranks, int32_t * split)
{
  if (!count || !split || n_cells < 0 || nranks < 1)
    return gfsb200_fail (GFSB200_ERR_ARG, "comm_splitters: bad argument");
  int64_t total = 0;
  for (int32_t i = 0; i < n_cells; i++) total += count[i];
  /* slice r starts at the first cell at which the running count reaches r*total/R: equal shares
     up to the population of one cell, and a cell never straddles two ranks */
  split[0] = 0;
  int64_t run = 0;
  int32_t cell = 0;
  for (int r = 1; r < nranks; r++) {
    const int64_t target = total*r/nranks;
    while (cell < n_cells && run + count[cell] <= target) run += count[cell++];
    split[r] = cell;
  }
  split[nranks] = n_cells;
  return GFSB200_OK;
}

/* Slices of the DEPTH-FIRST leaf order (adaptive trees).  The flat tree is level-ordered: a particle that
 * crosses from a leaf into a neighbour of another level lands far away in the cell index, i.e. -- with
 * slices of that index -- in another rank's slice (6 % of the C5 cloud after 20 steps).  In depth-first
 * order neighbouring leaves of any level stay close.  child0[c]: first child of cell c (children are
 * consecutive), < 0 for a leaf; the roots are cells 0 .. n_roots - 1; count[c]: particles in cell c.
 * owner[c] = rank of leaf c, 255 for a cell that is not a leaf.  Equal shares as gfsb200_comm_splitters
 * cuts them, along the depth-first order.  Within one level the owners are non-decreasing in the cell
 * index (the level's cells are Morton-ordered), so a rank's leaves of one level sit in one range of cells
 * that holds no other rank's leaf. */
static int owner_slices (const int32_t * child0, int32_t n_cells, int32_t n_roots, int dim,
			 const uint32_t * count, int nranks, const double * cum, uint8_t * owner);
static void splitters_by_share (const uint32_t * count, int32_t n_cells, int nranks, const double * cum, int32_t * split);

extern "C" int gfsb200_comm_owner_slices (const int32_t * child0, int32_t n_cells, int32_t n_roots, int dim,
					  const uint32_t * count, int nranks, uint8_t * owner)
{
  return owner_slices (child0, n_cells, n_roots, dim, count, nranks, NULL, owner);
}

static int owner_slices (const int32_t * child0, int32_t n_cells, int32_t n_roots, int dim,
			 const uint32_t * count, int nranks, const double * cum, uint8_t * owner)
{
  if (!child0 || !count || !owner || n_cells < 0 || n_roots < 0 || n_roots > n_cells || nranks < 1 || nranks > 254 ||
      (dim != 2 && dim != 3))
    return gfsb200_fail (GFSB200_ERR_ARG, "comm_owner_slices: bad argument");
  const int nc = 1 << dim;
  std::vector<int32_t> leaves, stack;
  leaves.reserve ((size_t) n_cells);
  for (int32_t r = 0; r < n_roots; r++) {
    stack.push_back (r);
    while (!stack.empty ()) {
      const int32_t c = stack.back ();
      stack.pop_back ();
      const int32_t c0 = child0[c];
      if (c0 < 0) { leaves.push_back (c); continue; }
      if (c0 + nc > n_cells)
	return gfsb200_fail (GFSB200_ERR_ARG, "comm_owner_slices: child index out of range");
      for (int k = nc - 1; k >= 0; k--) stack.push_back (c0 + k);
    }
  }
  std::vector<uint32_t> cp (leaves.size ());
  for (size_t i = 0; i < leaves.size (); i++) cp[i] = count[leaves[i]];
  int32_t split[256];
  if (cum)
    splitters_by_share (cp.data (), (int32_t) leaves.size (), nranks, cum, split);
  else {
    int r = gfsb200_comm_splitters (cp.data (), (int32_t) leaves.size (), nranks, split);
    if (r) return r;
  }
  memset (owner, 255, (size_t) n_cells);
  int q = 0;
  for (size_t i = 0; i < leaves.size (); i++) {
    while (q + 1 < nranks && (int32_t) i >= split[q + 1]) q++;
    owner[leaves[i]] = (uint8_t) q;
  }
  return GFSB200_OK;
}

/* owner of every cell after the last gfsb200_comm_rebalance (255: not a leaf / no owner) */
extern "C" int gfsb200_comm_owner_table (const gfsb200_comm * m, uint8_t * owner)
{
  if (!m || !owner) return gfsb200_fail (GFSB200_ERR_ARG, "comm_owner_table: bad argument");
  if (!m->owner_valid) return gfsb200_fail (GFSB200_ERR_STATE, "comm_owner_table: no rebalance since the tree was uploaded");
  const int32_t n = m->c->T.n_cells;
  if (m->by_owner)
    memcpy (owner, m->owner_of.data (), (size_t) n);
  else
    for (int q = 0; q < m->nranks; q++)
      for (int32_t c = m->split[q]; c < m->split[q + 1]; c++) owner[c] = (uint8_t) q;
  return GFSB200_OK;
}

/* the same with unequal target shares: slice r starts where the running count reaches cum[r]*total
 * (cum[0] = 0 <= cum[1] <= ... <= cum[nranks] = 1) */
static void splitters_by_share (const uint32_t * count, int32_t n_cells, int nranks, const double * cum, int32_t * split)
{
  int64_t total = 0;
  for (int32_t i = 0; i < n_cells; i++) total += count[i];
  split[0] = 0;
  int64_t run = 0;
  int32_t cell = 0;
  for (int r = 1; r < nranks; r++) {
    const int64_t target = (int64_t) ((double) total*cum[r]);
    while (cell < n_cells && run + count[cell] <= target) run += count[cell++];
    split[r] = cell;
  }
  split[nranks] = n_cells;
}

/* Target shares of the particles for the NEXT gfsb200_comm_rebalance (share[r] > 0, any scale; NULL:
 * equal shares again).  Equal numbers of particles are not equal work on an adaptive tree -- a particle
 * in a deep leaf costs the kernels more than one in a shallow leaf --: a caller that has timed its ranks
 * hands each one a share inversely proportional to its measured time per particle. */
extern "C" int gfsb200_comm_set_shares (gfsb200_comm * const * local, int n_local, const double * share)
{
  int r = check_local (local, n_local, "comm_set_shares");
  if (r) return r;
  const int R = local[0]->nranks;
  double cum[GFSB200_MAX_RANKS + 1];
  cum[0] = 0.;
  if (share) {
    double sum = 0.;
    for (int q = 0; q < R; q++) {
      if (!(share[q] > 0.)) return gfsb200_fail (GFSB200_ERR_ARG, "comm_set_shares: share[%d] is not positive", q);
      sum += share[q];
    }
    double run = 0.;
    for (int q = 0; q < R; q++) { run += share[q]; cum[q + 1] = run/sum; }
    cum[R] = 1.;
  }
  for (int k = 0; k < n_local; k++) {
    local[k]->shares_set = share != NULL;
    if (share) memcpy (local[k]->share_cum, cum, sizeof cum);
  }
  return GFSB200_OK;
}

extern "C" int gfsb200_comm_split (const gfsb200_comm * m, int32_t * split)
{
  if (!m || !split) return gfsb200_fail (GFSB200_ERR_ARG, "comm_split: bad argument");
  if (!m->owner_valid) return gfsb200_fail (GFSB200_ERR_STATE, "comm_split: no rebalance since the tree was uploaded");
  if (m->by_owner)
    return gfsb200_fail (GFSB200_ERR_STATE, "comm_split: on this (adaptive) tree the slices are ranges of the "
			 "depth-first leaf order, not of the cell index: gfsb200_comm_owner_table");
  memcpy (split, m->split, (m->nranks + 1)*sizeof (int32_t));
  return GFSB200_OK;
}

extern "C" int gfsb200_comm_rebalance (gfsb200_comm * const * local, int n_local)
{
  int r = check_local (local, n_local, "comm_rebalance");
  if (r) return r;
  const int R = local[0]->nranks;
  Nccl * N = R > 1 ? nccl_api () : NULL;
  const int32_t n_cells = local[0]->c->T.n_cells;
  for (int k = 0; k < n_local; k++) {
    if (!local[k]->c->have_tree) return gfsb200_fail (GFSB200_ERR_STATE, "comm_rebalance: no tree resident");
    if (local[k]->dep_mode)
      return gfsb200_fail (GFSB200_ERR_STATE, "comm_rebalance: a deposit is waiting for its gfsb200_deposit_allreduce");
  }
  if ((r = sync_all (local, n_local))) return r;

  /* 1. local sort by cell, per-cell counts, global counts */
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    gfsb200_ctx * c = m->c;
    CK (cudaSetDevice (c->device));
    if ((r = gfsb200_internal_ensure_aux (c, c->n > 0 ? c->n : 1))) return r;
    if ((r = gfsb200_internal_sort (c))) return r;
    if (m->hist_cap < (int64_t) n_cells + 1) {
      cudaFree (m->d_hist); m->d_hist = NULL; m->hist_cap = 0;
      CK (cudaMalloc ((void **) &m->d_hist, ((size_t) n_cells + 1)*sizeof (uint32_t)));
      m->hist_cap = (int64_t) n_cells + 1;
    }
    CK (cudaMemsetAsync (m->d_hist, 0, ((size_t) n_cells + 1)*sizeof (uint32_t), c->stream));
    if (c->n > 0) {
      gfsb200_launch_counter += 1;
      cell_histogram_kernel<<<(unsigned) ((c->n + 255)/256), 256, 0, c->stream>>> (c->n, c->key2, m->d_hist);
      CK (cudaGetLastError ());
    }
  }
  if (R > 1) {
    NK (N->GroupStart ());
    for (int k = 0; k < n_local; k++)
      NK (N->AllReduce (local[k]->d_hist, local[k]->d_hist, (size_t) n_cells, ncclUint32, ncclSum,
			local[k]->nccl, local[k]->c->stream));
    NK (N->GroupEnd ());
  }
  std::vector<uint32_t> hist ((size_t) n_cells);
  {
    gfsb200_ctx * c = local[0]->c;
    CK (cudaSetDevice (c->device));
    CK (cudaMemcpyAsync (hist.data (), local[0]->d_hist, (size_t) n_cells*sizeof (uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK (cudaStreamSynchronize (c->stream));
  }
  int32_t split[GFSB200_MAX_RANKS + 1];
  const double * cum = local[0]->shares_set ? local[0]->share_cum : NULL;
  if (cum)
    splitters_by_share (hist.data (), n_cells, R, cum, split);
  else if ((r = gfsb200_comm_splitters (hist.data (), n_cells, R, split))) return r;
  /* adaptive trees, several ranks: slices of the depth-first leaf order instead (gfsb200_comm_owner_slices) */
  const bool by_owner = R > 1 && local[0]->c->T.lattice_n1 <= 0 && !getenv ("GFSB200_SLICES_BY_CELL");
  std::vector<uint8_t> owner_of;
  if (by_owner) {
    gfsb200_ctx * c = local[0]->c;
    std::vector<int32_t> child0 ((size_t) n_cells);
    CK (cudaSetDevice (c->device));
    CK (cudaMemcpyAsync (child0.data (), c->T.child0, (size_t) n_cells*sizeof (int32_t), cudaMemcpyDeviceToHost, c->stream));
    CK (cudaStreamSynchronize (c->stream));
    owner_of.resize ((size_t) n_cells);
    if ((r = owner_slices (child0.data (), n_cells, c->T.n_roots, c->T.dim, hist.data (), R, cum, owner_of.data ())))
      return r;
    for (int k = 0; k < n_local; k++) {
      gfsb200_comm * m = local[k];
      CK (cudaSetDevice (m->c->device));
      if (m->owner_cap < (int64_t) n_cells) {
	cudaFree (m->d_owner_of); m->d_owner_of = NULL; m->owner_cap = 0;
	CK (cudaMalloc ((void **) &m->d_owner_of, (size_t) n_cells));
	m->owner_cap = n_cells;
      }
      CK (cudaMemcpyAsync (m->d_owner_of, owner_of.data (), (size_t) n_cells, cudaMemcpyHostToDevice, m->c->stream));
      /* every rank's share becomes one contiguous stretch of the list (still cell-sorted inside) */
      if ((r = gfsb200_internal_sort_by_owner (m->c, m->d_owner_of, R))) return r;
    }
  }

  /* 2. where the slices begin in every rank's sorted list; who sends how much to whom
     (one rank: the sorted list already is the one slice) */
  if (R > 1) {
  std::vector<int32_t> pos ((size_t) n_local*(R + 2)), matrix ((size_t) R*R, 0);
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    gfsb200_ctx * c = m->c;
    CK (cudaSetDevice (c->device));
    uint32_t bound[GFSB200_MAX_RANKS + 2];
    for (int q = 0; q <= R; q++)       /* bound[R]: where the particles outside the domain begin */
      bound[q] = by_owner ? (uint32_t) q : (uint32_t) split[q];
    uint32_t * d_bound = (uint32_t *) m->d_small;
    int32_t * d_pos = m->d_small + GFSB200_MAX_RANKS + 2;
    CK (cudaMemcpyAsync (d_bound, bound, (R + 1)*sizeof (uint32_t), cudaMemcpyHostToDevice, c->stream));
    gfsb200_launch_counter += 1;
    lower_bound_kernel<<<1, 32, 0, c->stream>>> (c->n, c->key2, R + 1, d_bound, d_pos);
    CK (cudaGetLastError ());
    CK (cudaMemcpyAsync (&pos[(size_t) k*(R + 2)], d_pos, (R + 1)*sizeof (int32_t), cudaMemcpyDeviceToHost, c->stream));
    CK (cudaStreamSynchronize (c->stream));
    pos[(size_t) k*(R + 2) + R + 1] = (int32_t) c->n;
  }
  {
    /* count matrix: row = sender, column = receiver */
    for (int k = 0; k < n_local; k++) {
      gfsb200_comm * m = local[k];
      int32_t row[GFSB200_MAX_RANKS];
      for (int q = 0; q < R; q++) row[q] = pos[(size_t) k*(R + 2) + q + 1] - pos[(size_t) k*(R + 2) + q];
      CK (cudaSetDevice (m->c->device));
      int32_t * d_mat = m->d_small + 2*GFSB200_MAX_RANKS + 8;
      CK (cudaMemcpyAsync (d_mat + m->rank*R, row, R*sizeof (int32_t), cudaMemcpyHostToDevice, m->c->stream));
    }
    if (R > 1) {
      NK (N->GroupStart ());
      for (int k = 0; k < n_local; k++) {
	gfsb200_comm * m = local[k];
	int32_t * d_mat = m->d_small + 2*GFSB200_MAX_RANKS + 8;
	NK (N->AllGather (d_mat + m->rank*R, d_mat, (size_t) R, ncclInt32, m->nccl, m->c->stream));
      }
      NK (N->GroupEnd ());
    }
    gfsb200_comm * m = local[0];
    CK (cudaSetDevice (m->c->device));
    CK (cudaMemcpyAsync (matrix.data (), m->d_small + 2*GFSB200_MAX_RANKS + 8, (size_t) R*R*sizeof (int32_t),
			 cudaMemcpyDeviceToHost, m->c->stream));
    if ((r = sync_all (local, n_local))) return r;
  }

  /* 3. move the particles: rank q's share of my list goes to q, and the shares of the others come
     in behind each other in the alternate SoA buffer (then: one more local sort merges the runs) */
  std::vector<int64_t> new_n (n_local);
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    gfsb200_ctx * c = m->c;
    const int32_t * P = &pos[(size_t) k*(R + 2)];
    const int64_t outside = (int64_t) P[R + 1] - P[R];     /* stay where they are: the list event culls them */
    int64_t total = outside;
    for (int q = 0; q < R; q++) total += matrix[(size_t) q*R + m->rank];
    if (total > INT32_MAX) return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "comm_rebalance: more than 2^31 particles on a rank");
    new_n[k] = total;
    CK (cudaSetDevice (c->device));
    if ((r = gfsb200_internal_reserve (c, total > c->n ? total : c->n))) return r;
  }
  if (R > 1) NK (N->GroupStart ());
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    gfsb200_ctx * c = m->c;
    const int32_t * P = &pos[(size_t) k*(R + 2)];
    const int cur = c->cur, alt = 1 - c->cur;
    CK (cudaSetDevice (c->device));
    int64_t off = 0;
    for (int q = 0; q < R; q++) {
      const int64_t n_in = matrix[(size_t) q*R + m->rank], n_out = P[q + 1] - P[q];
      for (int col = 0; col <= NCOL; col++) {
	const bool is_id = col == NCOL;
	const void * src = is_id ? (const void *) (c->id[cur] + P[q]) : (const void *) (c->col[cur][col] + P[q]);
	void * dst = is_id ? (void *) (c->id[alt] + off) : (void *) (c->col[alt][col] + off);
	const ncclDataType_t ty = is_id ? ncclUint32 : ncclDouble;
	if (q == m->rank) {
	  if (n_in > 0)
	    CK (cudaMemcpyAsync (dst, src, (size_t) n_in*(is_id ? 4 : 8), cudaMemcpyDeviceToDevice, c->stream));
	}
	else {
	  if (n_out > 0) NK (N->Send (src, (size_t) n_out, ty, q, m->nccl, c->stream));
	  if (n_in > 0) NK (N->Recv (dst, (size_t) n_in, ty, q, m->nccl, c->stream));
	}
      }
      off += n_in;
    }
    const int64_t outside = (int64_t) P[R + 1] - P[R];
    if (outside > 0)
      for (int col = 0; col <= NCOL; col++) {
	if (col == NCOL)
	  CK (cudaMemcpyAsync (c->id[alt] + off, c->id[cur] + P[R], (size_t) outside*4, cudaMemcpyDeviceToDevice, c->stream));
	else
	  CK (cudaMemcpyAsync (c->col[alt][col] + off, c->col[cur][col] + P[R], (size_t) outside*8,
			       cudaMemcpyDeviceToDevice, c->stream));
      }
  }
  if (R > 1) NK (N->GroupEnd ());
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    gfsb200_ctx * c = m->c;
    CK (cudaSetDevice (c->device));
    c->cur = 1 - c->cur;
    c->n = new_n[k];
    c->forces_recorded = false;
    c->esc_armed = false;
    if ((r = gfsb200_internal_ensure_aux (c, c->n > 0 ? c->n : 1))) return r;
    if ((r = gfsb200_internal_sort (c))) return r;
  }
  }

  /* 4. ownership: peers mapped for the current buffers, both deposit buffers cleared, slices published */
  bool remap = false;
  for (int k = 0; k < n_local; k++)
    if (local[k]->peer_generation != local[k]->c->tree_generation || !local[k]->c->deposit_buf[1]) remap = true;
  if (remap && (r = exchange_peers (local, n_local))) return r;
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    gfsb200_ctx * c = m->c;
    CK (cudaSetDevice (c->device));
    memcpy (m->split, split, sizeof split);
    m->by_owner = by_owner;
    m->range_lo.clear (); m->range_hi.clear ();
    if (by_owner) {
      /* the ranges of cells that hold this rank's leaves: runs of the cell index without another rank's
	 leaf in them (cells that are no leaves, 255, receive no deposit and may ride along), clipped to
	 the leaves of the GfsBox trees */
      m->owner_of = owner_of;
      int32_t run_lo = -1, last = -1;
      for (int32_t cell = c->leaf_lo; cell < c->leaf_hi; cell++) {
	const uint8_t o = owner_of[(size_t) cell];
	if (o == (uint8_t) m->rank) { if (run_lo < 0) run_lo = cell; last = cell; }
	else if (o != 255 && run_lo >= 0) { m->range_lo.push_back (run_lo); m->range_hi.push_back (last + 1); run_lo = -1; }
      }
      if (run_lo >= 0) { m->range_lo.push_back (run_lo); m->range_hi.push_back (last + 1); }
    }
    DevOwners own[2];
    memset (own, 0, sizeof own);
    for (int b = 0; b < 2; b++) {
      own[b].n = R; own[b].self = m->rank;
      for (int q = 0; q <= R; q++) own[b].split[q] = split[q];
      for (int q = 0; q < R; q++) own[b].base[q] = m->peer_dep[b][q];
    }
    CK (cudaMemcpyAsync (m->d_owners, own, sizeof own, cudaMemcpyHostToDevice, c->stream));
    for (int b = 0; b < 2; b++)
      CK (cudaMemsetAsync (c->deposit_buf[b], 0, (size_t) c->deposit_count*sizeof (double), c->stream));
    CK (cudaStreamSynchronize (c->stream));
    m->ahead_ok[0] = m->ahead_ok[1] = true;
    m->x_pending[0] = m->x_pending[1] = false;
    m->barrier_pending = false;
    m->owner_valid = true;
  }
  /* nobody deposits into a peer's slice before that peer has cleared it */
  if (R > 1) {
    NK (N->GroupStart ());
    for (int k = 0; k < n_local; k++)
      NK (N->AllReduce (local[k]->d_small, local[k]->d_small, 1, ncclInt32, ncclMin, local[k]->nccl, local[k]->c->stream));
    NK (N->GroupEnd ());
  }
  return sync_all (local, n_local);
}

/* ------------------------------------------------------------------ */
/* deposit: policy of the target buffer                                 */

extern "C" int gfsb200_comm_prepare_deposit (gfsb200_comm * m, int what, bool local_only, DevDeposit * D)
{
  gfsb200_ctx * c = m->c;
  const int t = c->dep_which;
  const size_t n = c->T.n_cells;
  if (!c->deposit_buf[1]) {
    CK (cudaMalloc ((void **) &c->deposit_buf[1], (size_t) c->deposit_count*sizeof (double)));
    CK (cudaMemsetAsync (c->deposit_buf[1], 0, (size_t) c->deposit_count*sizeof (double), c->stream));
  }
  const bool owner_possible = m->owner_valid && (m->p2p || m->nranks == 1) &&
    m->peer_generation == c->tree_generation && m->exchange_mode == GFSB200_EXCHANGE_AUTO;
  const int mode = owner_possible && !local_only ? MODE_OWNER : MODE_LOCAL;
  if (local_only && m->nranks > 1 && m->exchange_mode != GFSB200_EXCHANGE_ALLREDUCE)
    return gfsb200_fail (GFSB200_ERR_STATE, "the smoothed deposit reaches cells of every rank: select "
			 "GFSB200_EXCHANGE_ALLREDUCE (gfsb200_comm_set_exchange) on a multi-GPU communicator");
  if (m->dep_mode && m->dep_mode != mode)
    return gfsb200_fail (GFSB200_ERR_STATE, "deposit: owner-slice and whole-buffer deposits cannot share a step");
  if (what & m->dep_done)
    return gfsb200_fail (GFSB200_ERR_STATE, "deposit: the same component was deposited twice without a "
			 "gfsb200_deposit_allreduce in between");
  /* the target's previous exchange (two steps ago) must have drained */
  if (m->x_pending[t]) {
    CK (cudaStreamWaitEvent (c->stream, m->ev_x[t], 0));
    m->x_pending[t] = false;
  }
  D->local = c->deposit_buf[t];
  if (mode == MODE_LOCAL) {
    if (what & 1) CK (cudaMemsetAsync (c->deposit_buf[t], 0, n*sizeof (double), c->stream));
    if (what & 2) CK (cudaMemsetAsync (c->deposit_buf[t] + n, 0, (size_t) c->T.dim*n*sizeof (double), c->stream));
    D->own_lo = 0; D->own_hi = (int32_t) n;
    D->peers = NULL;
    m->ahead_ok[t] = false;
  }
  else {
    if (!m->ahead_ok[t])
      return gfsb200_fail (GFSB200_ERR_STATE, "deposit: the rank's slice of the target buffer was not cleared ahead "
			   "(whole-buffer deposits were mixed in): call gfsb200_comm_rebalance");
    const int32_t lo = m->split[m->rank], hi = m->split[m->rank + 1];
    /* (the own slice of the OTHER buffer is cleared for the next step by gfsb200_deposit_allreduce,
       behind this step's kernels: see the ordering note at the top of the file)
       Nothing may be reduced into a peer's slice before that peer has cleared it: the deposit
       kernel waits for B(n-1) -- the barrier only, not the pushes behind it. */
    if (m->barrier_pending) {
      CK (cudaStreamWaitEvent (c->stream, m->ev_barrier, 0));
      m->barrier_pending = false;
    }
    D->own_lo = lo; D->own_hi = hi;
    D->peers = m->nranks > 1 ? m->d_owners + t : NULL;
    if (m->by_owner && m->nranks > 1) {
      D->owner_of = m->d_owner_of;
      D->self = m->rank;
    }
    D->local_gpu_scope = getenv ("GFSB200_LOCAL_RED_GPU_SCOPE") != NULL;
  }
  m->dep_mode = mode;
  m->dep_done |= what;
  return GFSB200_OK;
}

/* the part of the rank's slice a deposit can reach: slices partition ALL cells, but only leaves of
   the GfsBox trees receive anything -- nothing else needs clearing or pushing (on a uniform tree
   the first slice would otherwise carry every non-leaf level along) */
static void slice_leaves (const gfsb200_comm * m, int32_t * lo, int32_t * hi)
{
  *lo = m->split[m->rank] > m->c->leaf_lo ? m->split[m->rank] : m->c->leaf_lo;
  *hi = m->split[m->rank + 1] < m->c->leaf_hi ? m->split[m->rank + 1] : m->c->leaf_hi;
  if (*hi < *lo) *hi = *lo;
}

/* the same as a list of ranges: one (slice_leaves) when the slices are ranges of the cell index, the
   per-level runs of the owner table on adaptive trees */
static void slice_ranges (const gfsb200_comm * m, std::vector<int32_t> & lo, std::vector<int32_t> & hi)
{
  lo.clear (); hi.clear ();
  if (m->by_owner) { lo = m->range_lo; hi = m->range_hi; return; }
  int32_t a, b;
  slice_leaves (m, &a, &b);
  if (b > a) { lo.push_back (a); hi.push_back (b); }
}

static int stat_begin (gfsb200_comm * m)
{
  if (m->tev_used + 2 > 8192) return GFSB200_OK;
  if (m->tev_used + 2 > m->tev.size ()) {
    cudaEvent_t a, b;
    CK (cudaEventCreate (&a));
    CK (cudaEventCreate (&b));
    m->tev.push_back (a); m->tev.push_back (b);
  }
  CK (cudaEventRecord (m->tev[m->tev_used], m->stream));
  return GFSB200_OK;
}

static int stat_end (gfsb200_comm * m)
{
  if (m->tev_used + 2 > 8192 || m->tev_used + 2 > m->tev.size ()) return GFSB200_OK;
  CK (cudaEventRecord (m->tev[m->tev_used + 1], m->stream));
  m->tev_used += 2;
  return GFSB200_OK;
}

extern "C" int gfsb200_deposit_allreduce (gfsb200_comm * const * local, int n_local)
{
  int r = check_local (local, n_local, "deposit_allreduce");
  if (r) return r;
  const int R = local[0]->nranks;
  Nccl * N = R > 1 ? nccl_api () : NULL;
  const int mode = local[0]->dep_mode;
  for (int k = 0; k < n_local; k++) {
    if (!local[k]->dep_mode)
      return gfsb200_fail (GFSB200_ERR_STATE, "deposit_allreduce: nothing has been deposited since the last exchange");
    if (local[k]->dep_mode != mode)
      return gfsb200_fail (GFSB200_ERR_STATE, "deposit_allreduce: the ranks of this process deposited in different modes");
  }
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    gfsb200_ctx * c = m->c;
    CK (cudaSetDevice (c->device));
    CK (cudaEventRecord (m->ev_dep, c->stream));
    CK (cudaStreamWaitEvent (m->stream, m->ev_dep, 0));
    if (mode == MODE_OWNER) {
      /* Z(n+1): clear my slice of the OTHER buffer for the next step -- on the communication
	 stream, behind this step's kernels and the previous exchange of that buffer (same stream),
	 before my arrival at B(n).  The context's stream goes straight on to the next step's cell
	 pass; its next deposit kernel waits for ev_barrier (gfsb200_comm_prepare_deposit). */
      const int o = 1 - c->dep_which;
      const size_t n = c->T.n_cells;
      std::vector<int32_t> lo, hi;
      slice_ranges (m, lo, hi);
      for (size_t g = 0; g < lo.size (); g++)
	CK (cudaMemset2DAsync (c->deposit_buf[o] + lo[g], n*sizeof (double), 0, (size_t) (hi[g] - lo[g])*sizeof (double),
			       (size_t) c->T.dim + 1, m->stream));
      m->ahead_ok[o] = true;
      if (R == 1) {
	CK (cudaEventRecord (m->ev_barrier, m->stream));
	m->barrier_pending = true;
      }
    }
    if (!(mode == MODE_OWNER && R > 1) && (r = stat_begin (m))) return r;
    m->epoch++;
  }
  if (mode == MODE_LOCAL) {
    if (R > 1) {
      NK (N->GroupStart ());
      for (int k = 0; k < n_local; k++) {
	gfsb200_comm * m = local[k];
	double * buf = m->c->deposit_buf[m->c->dep_which];
	NK (N->AllReduce (buf, buf, (size_t) m->c->deposit_count, ncclDouble, ncclSum, m->nccl, m->stream));
	m->bytes_sent = (int64_t) (2.*(R - 1)/R*m->c->deposit_count*sizeof (double));
      }
      NK (N->GroupEnd ());
    }
  }
  else if (R > 1) {
    for (int k = 0; k < n_local; k++) {
      gfsb200_comm * m = local[k];
      gfsb200_ctx * c = m->c;
      const int t = c->dep_which;
      const size_t n = c->T.n_cells;
      std::vector<int32_t> lo, hi;
      slice_ranges (m, lo, hi);
      CK (cudaSetDevice (c->device));
      PeerFlags pf;
      for (int q = 0; q < GFSB200_MAX_RANKS; q++) pf.p[q] = q < R ? m->peer_flags[q] : NULL;
      /* B(n): every rank has finished its deposit kernel, so every reduction into my slice is in */
      gfsb200_launch_counter += 1;
      comm_flags_kernel<<<1, 32, 0, m->stream>>> (pf, m->flags, m->rank, R, FLAG_ARRIVE, m->epoch, 1, 1);
      CK (cudaGetLastError ());
      CK (cudaEventRecord (m->ev_barrier, m->stream));
      m->barrier_pending = true;
      if ((r = stat_begin (m))) return r;          /* the transfer is timed from here: the barrier is rank skew */
      /* my slice to every peer: copy engines over NVLink, no SM time.  One contiguous copy per
	 peer and component (a strided cudaMemcpy2DAsync between peers measured 100 GB/s against
	 320+ for plain copies), spread over a few streams so that several engines and links work
	 at once. */
      int64_t slice_cells = 0;
      for (size_t g = 0; g < lo.size (); g++) slice_cells += hi[g] - lo[g];
      if (slice_cells > 0) {
	CK (cudaEventRecord (m->ev_fork, m->stream));
	for (int j = 0; j < N_PUSH; j++)
	  CK (cudaStreamWaitEvent (m->push[j], m->ev_fork, 0));
	int k2 = 0;
	for (int d = 1; d < R; d++) {
	  const int q = (m->rank + d) % R;
	  for (int comp = 0; comp <= c->T.dim; comp++)
	    for (size_t g = 0; g < lo.size (); g++, k2++)
	      CK (cudaMemcpyAsync (m->peer_dep[t][q] + comp*n + lo[g], c->deposit_buf[t] + comp*n + lo[g],
				   (size_t) (hi[g] - lo[g])*sizeof (double), cudaMemcpyDefault, m->push[k2 % N_PUSH]));
	}
	for (int j = 0; j < N_PUSH; j++) {
	  CK (cudaEventRecord (m->ev_join[j], m->push[j]));
	  CK (cudaStreamWaitEvent (m->stream, m->ev_join[j], 0));
	}
      }
      /* DONE(n) to every rank, behind the pushes: 4-byte copies by the copy engine, not a kernel --
	 a kernel would queue behind the next step's persistent step kernel, which fills every SM */
      for (int q = 0; q < R; q++)
	CK (cudaMemcpyAsync (m->peer_flags[q] + FLAG_DONE*GFSB200_MAX_RANKS + m->rank,
			     m->flags + FLAG_SRC*GFSB200_MAX_RANKS, sizeof (uint32_t), cudaMemcpyDefault, m->stream));
      m->bytes_sent = (int64_t) (R - 1)*(c->T.dim + 1)*slice_cells*(int64_t) sizeof (double);
    }
  }
  for (int k = 0; k < n_local; k++) {
    gfsb200_comm * m = local[k];
    gfsb200_ctx * c = m->c;
    const int t = c->dep_which;
    CK (cudaSetDevice (c->device));
    if ((r = stat_end (m))) return r;
    CK (cudaEventRecord (m->ev_x[t], m->stream));
    m->x_pending[t] = true;
    m->x_mode[t] = mode;
    m->x_epoch[t] = m->epoch;
    m->ahead_ok[t] = false;
    m->dep_mode = m->dep_done = 0;
    c->dep_result = t;
    c->dep_which = 1 - t;
    c->deposit = c->deposit_buf[1 - t];
  }
  return GFSB200_OK;
}

extern "C" int gfsb200_deposit_wait (gfsb200_comm * m)
{
  if (!m || !m->c) return gfsb200_fail (GFSB200_ERR_ARG, "deposit_wait: null communicator");
  gfsb200_ctx * c = m->c;
  const int t = c->dep_result;
  CK (cudaSetDevice (c->device));
  if (!m->x_pending[t]) return GFSB200_OK;
  CK (cudaStreamWaitEvent (c->stream, m->ev_x[t], 0));
  if (m->x_mode[t] == MODE_OWNER && m->nranks > 1) {
    /* ... and the slices of the others have landed in my copy */
    PeerFlags pf;
    for (int q = 0; q < GFSB200_MAX_RANKS; q++) pf.p[q] = q < m->nranks ? m->peer_flags[q] : NULL;
    gfsb200_launch_counter += 1;
    comm_flags_kernel<<<1, 32, 0, c->stream>>> (pf, m->flags, m->rank, m->nranks, FLAG_DONE, m->x_epoch[t], 0, 1);
    CK (cudaGetLastError ());
  }
  return GFSB200_OK;
}

extern "C" int gfsb200_comm_exchange_stats (gfsb200_comm * m, double * ms, int64_t * n, int64_t * bytes_sent)
{
  if (!m || !m->c) return gfsb200_fail (GFSB200_ERR_ARG, "comm_exchange_stats: null communicator");
  CK (cudaSetDevice (m->c->device));
  CK (cudaStreamSynchronize (m->stream));
  double total = 0.;
  for (size_t i = 0; i + 1 < m->tev_used; i += 2) {
    float t = 0.f;
    CK (cudaEventElapsedTime (&t, m->tev[i], m->tev[i + 1]));
    total += t;
  }
  const int64_t cnt = (int64_t) (m->tev_used/2);
  if (ms) *ms = cnt ? total/cnt : 0.;
  if (n) *n = cnt;
  if (bytes_sent) *bytes_sent = m->bytes_sent;
  m->tev_used = 0;
  return GFSB200_OK;
}
