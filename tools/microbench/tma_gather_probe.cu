/* Micro-benchmark (not product code): what does the vertex gather of a lattice tree cost
 *  GATHER 1: per-lane LDG.128 + LDG.64 from the split table (what step_kernel_wpipe does)
 *  GATHER 2: one TMA tensor load per aligned 2x2x2 block of cells -- the 3x3x3 vertices of the block,
 *            864 B -- as box (4,3,3,3) of a (4, n1, n1, n1) fp64 tensor, then LDS
 *  GATHER 3: the same bytes as box (12,3,3) of a (4 n1, n1, n1) tensor (9 segments of 96 B)
 * alone and together with the particle stream of the step kernel (STREAM: 8 bulk copies of 256 B per
 * 32-particle tile, 11 LDS.64, 6 STG.64).  Synthetic access pattern of C2: sorted particles, ppc per leaf,
 * tiles handed to warps round-robin, 7 CTAs x 4 warps per SM.
 *   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tma_gather_probe tma_gather_probe.cu
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf (stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString (e_)); exit (1); } } while (0)

constexpr int LV = 7, N = 1 << LV, N1 = N + 1;
constexpr int WARPS = 4;
constexpr int BOX_BYTES = 4*27*8;       /* 864 */
constexpr int SLOT_BYTES = 896;         /* 128-byte aligned slots */
constexpr int MAXB = 3;
constexpr int STREAM_BYTES = 8*32*8;    /* one staged tile */

__device__ __forceinline__ unsigned compact3 (unsigned v)
{
  v &= 0x09249249;
  v = (v | (v >> 2)) & 0x030c30c3;
  v = (v | (v >> 4)) & 0x0300f00f;
  v = (v | (v >> 8)) & 0x030000ff;
  v = (v | (v >> 16)) & 0x3ff;
  return v;
}

__device__ __forceinline__ uint32_t smem_u32 (const void * p) { return (uint32_t) __cvta_generic_to_shared (p); }
__device__ __forceinline__ void mbar_init (uint64_t * bar, unsigned count)
{ asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx (uint64_t * bar, unsigned bytes)
{ asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait (uint64_t * bar, unsigned parity)
{
  asm volatile ("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
		:: "r"(smem_u32 (bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one ()
{
  unsigned pred;
  asm volatile ("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_async_shared () { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tensor4_g2s (void * dst, const CUtensorMap * map, int x, int y, int z, uint64_t * bar)
{
  asm volatile ("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
		"[%0], [%1, {%2, %3, %4, %5}], [%6];"
		:: "r"(smem_u32 (dst)), "l"(map), "r"(0), "r"(x), "r"(y), "r"(z), "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ void tensor3_g2s (void * dst, const CUtensorMap * map, int x, int y, int z, uint64_t * bar)
{
  asm volatile ("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
		"[%0], [%1, {%2, %3, %4}], [%5];"
		:: "r"(smem_u32 (dst)), "l"(map), "r"(4*x), "r"(y), "r"(z), "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s (void * dst, const void * src, unsigned bytes, uint64_t * bar)
{
  asm volatile ("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32 (dst)), "l"(src), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}

/* the particle -> leaf map of the synthetic cloud */
__device__ __forceinline__ void cell_of (int64_t p, double ppc_inv, int & kx, int & ky, int & kz)
{
  unsigned c = (unsigned) ((double) p*ppc_inv);
  c &= (1u << (3*LV)) - 1;
  kx = compact3 (c); ky = compact3 (c >> 1); kz = compact3 (c >> 2);
}

struct Cols { const double * in[8]; double * out[6]; };

template <int GATHER, bool STREAM>
__global__ void __launch_bounds__(32*WARPS, 7)
probe_kernel (const __grid_constant__ CUtensorMap map, const double2 * __restrict__ uv, const double * __restrict__ w,
	      Cols C, int n_tiles, double ppc_inv, double * __restrict__ out, unsigned long long * __restrict__ fallback_tiles)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t gbar[WARPS], sbar[WARPS][2];
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync (0xffffffffu, (int) (threadIdx.x >> 5), 0);
  const int first = blockIdx.x*WARPS + warp, stride = gridDim.x*WARPS;
  constexpr int PER_WARP = (GATHER >= 2 ? MAXB*SLOT_BYTES : 0) + (STREAM ? 2*STREAM_BYTES : 0);
  unsigned char * gbuf = smem_raw + (size_t) warp*PER_WARP;
  double (* sbuf)[8][32] = reinterpret_cast<double (*)[8][32]> (gbuf + (GATHER >= 2 ? MAXB*SLOT_BYTES : 0));

  auto issue_stream = [&] (int s, int tile) {
    mbar_expect_tx (&sbar[warp][s], STREAM_BYTES);
#pragma unroll
    for (int c = 0; c < 8; c++)
      bulk_g2s (&sbuf[s][c][0], C.in[c] + (int64_t) tile*32, 256, &sbar[warp][s]);
  };
  if (elect_one ()) {
    mbar_init (&gbar[warp], 1);
    mbar_init (&sbar[warp][0], 1); mbar_init (&sbar[warp][1], 1);
    fence_async_shared ();
    if (STREAM)
      for (int s = 0; s < 2; s++)
	if (first + s*stride < n_tiles) issue_stream (s, first + s*stride);
  }
  __syncwarp ();

  double acc = 0.;
  int s = 0;
  unsigned sparity = 0, gparity = 0;
  for (int tile = first; tile < n_tiles; tile += stride) {
    double px = 0., py = 0., pz = 0.;
    if (STREAM) {
      mbar_wait (&sbar[warp][s], sparity);
      px = sbuf[s][0][lane]; py = sbuf[s][1][lane]; pz = sbuf[s][2][lane];
    }
    int kx, ky, kz;
    cell_of ((int64_t) tile*32 + lane, ppc_inv, kx, ky, kz);
    const double tx = 0.25 + 0.01*lane + px, ty = 0.5 + py, tz = 0.75 + pz;
    double su = 0., sv = 0., sw = 0.;
    if (GATHER == 1) {
      const int b = (kz*N1 + ky)*N1 + kx;
#pragma unroll
      for (int plane = 0; plane < 2; plane++) {
	double fu[4], fv[4], fw[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
	  const int id = b + plane*N1*N1 + (k >> 1)*N1 + (k & 1);
	  const double2 ab = __ldg (uv + id);
	  fu[k] = ab.x; fv[k] = ab.y; fw[k] = __ldg (w + id);
	}
	const double wz = plane ? tz : 1. - tz;
	su += wz*(fma (tx, fu[1] - fu[0], fu[0])*(1. - ty) + fma (tx, fu[3] - fu[2], fu[2])*ty);
	sv += wz*(fma (tx, fv[1] - fv[0], fv[0])*(1. - ty) + fma (tx, fv[3] - fv[2], fv[2])*ty);
	sw += wz*(fma (tx, fw[1] - fw[0], fw[0])*(1. - ty) + fma (tx, fw[3] - fw[2], fw[2])*ty);
      }
    }
    else if (GATHER >= 2) {
      const int blk = ((kz >> 1) << 12) | ((ky >> 1) << 6) | (kx >> 1);
      const int prev = __shfl_up_sync (0xffffffffu, blk, 1);
      const bool head = lane == 0 || prev != blk;
      const unsigned heads = __ballot_sync (0xffffffffu, head);
      const int nblk = __popc (heads);
      const int slot = __popc (heads & (0xffffffffu >> (31 - lane))) - 1;
      if (nblk <= MAXB) {
	unsigned h = heads;
	int bx[MAXB], by[MAXB], bz[MAXB];
#pragma unroll
	for (int j = 0; j < MAXB; j++) {
	  const int src = h ? __ffs (h) - 1 : 0;
	  h &= h - 1;
	  const int bj = __shfl_sync (0xffffffffu, blk, src);
	  bx[j] = (bj & 63) << 1; by[j] = ((bj >> 6) & 63) << 1; bz[j] = (bj >> 12) << 1;
	}
	if (elect_one ()) {
	  mbar_expect_tx (&gbar[warp], nblk*BOX_BYTES);
#pragma unroll
	  for (int j = 0; j < MAXB; j++)
	    if (j < nblk) {
	      if (GATHER == 2) tensor4_g2s (gbuf + j*SLOT_BYTES, &map, bx[j], by[j], bz[j], &gbar[warp]);
	      else             tensor3_g2s (gbuf + j*SLOT_BYTES, &map, bx[j], by[j], bz[j], &gbar[warp]);
	    }
	}
	__syncwarp ();
	mbar_wait (&gbar[warp], gparity); gparity ^= 1;
	const int cx = kx & 1, cy = ky & 1, cz = kz & 1;
	const unsigned char * base = gbuf + slot*SLOT_BYTES;
#pragma unroll
	for (int plane = 0; plane < 2; plane++) {
	  double fu[4], fv[4], fw[4];
#pragma unroll
	  for (int k = 0; k < 4; k++) {
	    const int r = ((cz + plane)*3 + cy + (k >> 1))*3 + cx + (k & 1);
	    const double2 ab = *reinterpret_cast<const double2 *> (base + r*32);
	    fu[k] = ab.x; fv[k] = ab.y;
	    fw[k] = *reinterpret_cast<const double *> (base + r*32 + 16);
	  }
	  const double wz = plane ? tz : 1. - tz;
	  su += wz*(fma (tx, fu[1] - fu[0], fu[0])*(1. - ty) + fma (tx, fu[3] - fu[2], fu[2])*ty);
	  sv += wz*(fma (tx, fv[1] - fv[0], fv[0])*(1. - ty) + fma (tx, fv[3] - fv[2], fv[2])*ty);
	  sw += wz*(fma (tx, fw[1] - fw[0], fw[0])*(1. - ty) + fma (tx, fw[3] - fw[2], fw[2])*ty);
	}
      }
      else if (lane == 0)
	atomicAdd (fallback_tiles, 1ULL);
    }
    acc += su + 2.*sv + 3.*sw;
    if (STREAM) {
      /* late fetch + re-read of the position, as in step_kernel_wpipe, then the six stores */
      const double vx = sbuf[s][3][lane], vy = sbuf[s][4][lane], vz = sbuf[s][5][lane];
      const double m = sbuf[s][6][lane], vol = sbuf[s][7][lane];
      const double x2 = sbuf[s][0][lane] + su*1e-30, y2 = sbuf[s][1][lane] + sv*1e-30, z2 = sbuf[s][2][lane] + sw*1e-30;
      const int64_t i = (int64_t) tile*32 + lane;
      __stcs (C.out[0] + i, x2 + vx*m); __stcs (C.out[1] + i, y2 + vy*m); __stcs (C.out[2] + i, z2 + vz*m);
      __stcs (C.out[3] + i, vx + vol); __stcs (C.out[4] + i, vy + vol); __stcs (C.out[5] + i, vz + vol);
    }
    __syncwarp ();
    if (STREAM || GATHER >= 2) {
      if (elect_one ()) {
	fence_async_shared ();
	if (STREAM && tile + 2*stride < n_tiles) issue_stream (s, tile + 2*stride);
      }
      if (STREAM) { if (++s == 2) { s = 0; sparity ^= 1; } }
    }
  }
  out[blockIdx.x*blockDim.x + threadIdx.x] = acc;
}

typedef CUresult (*EncodeFn) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
			      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
			      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main (int argc, char ** argv)
{
  const double ppc = argc > 1 ? atof (argv[1]) : 4.77;
  const int64_t n_particles = (int64_t) ((double) (1 << (3*LV))*ppc) & ~255LL;
  const int n_tiles = (int) (n_particles/32);
  const int64_t nv = (int64_t) N1*N1*N1;
  std::vector<double> h4 (4*nv), huv (2*nv), hw (nv);
  for (int64_t i = 0; i < nv; i++) {
    const double u = 1e-3*(double) (i % 1000), v = 2e-3*(double) (i % 777), w = 3e-3*(double) (i % 555);
    h4[4*i] = u; h4[4*i + 1] = v; h4[4*i + 2] = w; h4[4*i + 3] = 0.;
    huv[2*i] = u; huv[2*i + 1] = v; hw[i] = w;
  }
  double * d4, * duv, * dw, * out;
  unsigned long long * fb;
  CK (cudaMalloc (&d4, 4*nv*8)); CK (cudaMalloc (&duv, 2*nv*8)); CK (cudaMalloc (&dw, nv*8));
  CK (cudaMemcpy (d4, h4.data (), 4*nv*8, cudaMemcpyHostToDevice));
  CK (cudaMemcpy (duv, huv.data (), 2*nv*8, cudaMemcpyHostToDevice));
  CK (cudaMemcpy (dw, hw.data (), nv*8, cudaMemcpyHostToDevice));
  const int grid = 148*7, threads = 32*WARPS;
  CK (cudaMalloc (&out, (size_t) grid*threads*8));
  CK (cudaMalloc (&fb, 8)); CK (cudaMemset (fb, 0, 8));
  Cols C;
  for (int c = 0; c < 8; c++) { double * p; CK (cudaMalloc (&p, n_particles*8)); CK (cudaMemset (p, 0, n_particles*8)); C.in[c] = p; }
  for (int c = 0; c < 6; c++) { CK (cudaMalloc (&C.out[c], n_particles*8)); }

  EncodeFn encode = NULL;
  cudaDriverEntryPointQueryResult qres;
  CK (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", (void **) &encode, cudaEnableDefault, &qres));
  if (!encode) { fprintf (stderr, "no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap map4, map3;
  {
    const cuuint64_t dims[4] = { 4, (cuuint64_t) N1, (cuuint64_t) N1, (cuuint64_t) N1 };
    const cuuint64_t strides[3] = { 32, 32ull*N1, 32ull*N1*N1 };
    const cuuint32_t box[4] = { 4, 3, 3, 3 }, estr[4] = { 1, 1, 1, 1 };
    CUresult r = encode (&map4, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d4, dims, strides, box, estr,
			 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
			 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf (stderr, "encode 4d failed: %d\n", (int) r); return 1; }
  }
  {
    const cuuint64_t dims[3] = { 4ull*N1, (cuuint64_t) N1, (cuuint64_t) N1 };
    const cuuint64_t strides[2] = { 32ull*N1, 32ull*N1*N1 };
    const cuuint32_t box[3] = { 12, 3, 3 }, estr[3] = { 1, 1, 1 };
    CUresult r = encode (&map3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d4, dims, strides, box, estr,
			 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
			 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf (stderr, "encode 3d failed: %d\n", (int) r); return 1; }
  }

  cudaEvent_t e0, e1;
  CK (cudaEventCreate (&e0)); CK (cudaEventCreate (&e1));
  int clock_khz; CK (cudaDeviceGetAttribute (&clock_khz, cudaDevAttrClockRate, 0));
  std::vector<double> h (grid*threads);
  const int reps = 10;
  auto run = [&] (const char * name, auto kern, const CUtensorMap & map, size_t smem) {
    CK (cudaFuncSetAttribute (kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    float ms = 0.;
    for (int pass = 0; pass < 2; pass++) {
      CK (cudaEventRecord (e0));
      for (int k = 0; k < reps; k++)
	kern<<<grid, threads, smem>>> (map, (const double2 *) duv, dw, C, n_tiles, 1./ppc, out, fb);
      CK (cudaEventRecord (e1)); CK (cudaEventSynchronize (e1)); CK (cudaGetLastError ());
      CK (cudaEventElapsedTime (&ms, e0, e1));
    }
    ms /= reps;
    CK (cudaMemcpy (h.data (), out, h.size ()*8, cudaMemcpyDeviceToHost));
    double sum = 0.; for (double v : h) sum += v;
    const double cyc = ms*1e-3*clock_khz*1e3/((double) n_tiles/148.);
    printf ("%-44s %8.4f ms  %7.1f cycles per tile and SM  checksum %.12e\n", name, ms, cyc, sum);
    fflush (stdout);
  };
  const size_t G = (size_t) WARPS*MAXB*SLOT_BYTES, S = (size_t) WARPS*2*STREAM_BYTES;
  run ("gather LDG.128+LDG.64", probe_kernel<1, false>, map4, 0);
  run ("gather TMA box (4,3,3,3)", probe_kernel<2, false>, map4, G);
  run ("gather TMA box (12,3,3)", probe_kernel<3, false>, map3, G);
  run ("stream only (8 bulk in, 11 LDS, 6 STG)", probe_kernel<0, true>, map4, S);
  run ("stream + gather LDG", probe_kernel<1, true>, map4, S);
  run ("stream + gather TMA box (4,3,3,3)", probe_kernel<2, true>, map4, G + S);
  run ("stream + gather TMA box (12,3,3)", probe_kernel<3, true>, map3, G + S);
  unsigned long long hfb; CK (cudaMemcpy (&hfb, fb, 8, cudaMemcpyDeviceToHost));
  printf ("tiles %d, particles per leaf %.2f, tiles with more than %d blocks: %llu\n", n_tiles, ppc, MAXB, hfb);
  return 0;
}
