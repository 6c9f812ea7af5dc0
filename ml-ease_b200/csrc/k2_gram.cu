// k2_gram.cu -- K2: weighted Gram  G = Xt^T Xt  (Xt = diag(sqrt d) X in bf16, written by K1), the
// data term of LogisticRegressionL2.hessian (llf/LogisticRegressionL2.java:258-297):
//     H[m][n] = (m==n ? 1/priorVar[m] : 0) + sum_i D_ii x_im x_in ,  D_ii = w_i p_i (1-p_i).
// It genuinely is a dense GEMM (K = rows, M = N = features), so it runs on the 5th-gen tensor
// cores: TMA (tensor map, 128B swizzle) -> shared memory -> tcgen05.mma (bf16 x bf16 -> fp32 in
// TMEM) -> tcgen05.ld epilogue.  Both operands are tiles of the SAME row-major matrix, i.e. they are
// MN-major ("transposed") UMMA operands: no transpose pass over X is ever made.
//
// Work decomposition: output tiles of 128 (M) x 256 (N) restricted to the lower block triangle,
// split-K over row slices; each CTA owns one (tile, slice), accumulates it in TMEM (256 columns)
// and stores the fp32 partial to Hpart[slice] (plain stores, deterministic).  chol_prep_kernel
// (k3_cholesky.cu) sums the slices in fixed order and adds diag(q).
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one
// elected lane) + TMEM allocator, warps 2-5 = epilogue (tcgen05.ld 32x32b, one TMEM lane quadrant
// each: warp_id % 4).
//
// A fp32 SIMT kernel computing the same partials from the same bf16 operand is kept ONLY as a
// debug cross-check reachable through mlease_objective(tensor=0); the product path never uses it.
#include <cuda.h>

#include "kernels.cuh"

namespace mlease {

// ------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------
constexpr int GM = 128;          // tile rows  (UMMA M)
constexpr int GN = 256;          // tile cols  (UMMA N)
constexpr int GK = 64;           // K (data rows) per pipeline stage
constexpr int UK = 16;           // K per tcgen05.mma (bf16)
constexpr int GSTAGES = 4;
constexpr int G_A_BYTES = GK * GM * 2;   // 16 KB : 2 boxes of [64 k][64 feat]
constexpr int G_B_BYTES = GK * GN * 2;   // 32 KB : 4 boxes
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_BOX_BYTES = GK * 64 * 2; // 8 KB, one TMA box = 64 k-rows x 128 B
constexpr int G_THREADS = 192;
constexpr size_t G_SMEM = (size_t)GSTAGES * G_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

// UMMA shared-memory descriptor, MN-major operand, SWIZZLE_128B (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128)
// canonical MN-major SW128 layout (bf16): ((64 elems,m),(8,k)) : ((1,LBO),(128B,SBO)):
//   SBO = 1024 B between 8-row K groups, LBO = GK*128 B between 64-element MN groups (one TMA box).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b_format BF16 (1) @7/@10,
// a_major/b_major = MN (1) @15/@16, N>>3 @17, M>>4 @24.
__device__ __forceinline__ uint32_t umma_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct GramTile { short bi, bj; };   // 128-row block index, 256-col block index

__global__ void __launch_bounds__(G_THREADS, 1)
gram_tcgen05_kernel(const Problem* __restrict__ probs, const CUtensorMap* __restrict__ tmaps, const GramTile* __restrict__ tiles,
                    int ntiles, int force) {
  const int pidx = blockIdx.z;
  const Problem& pb = probs[pidx];
  Ctrl* ctrl = pb.ctrl;
  if (!force && (ctrl->done || !ctrl->need_hess)) return;
  const CUtensorMap* tmap = &tmaps[pidx];
  const GramTile tile = tiles[blockIdx.x];
  const int slice = blockIdx.y, nslices = gridDim.y;
  const int Dp = pb.Dp;
  const long long ksteps_total = (pb.n + GK - 1) / GK;
  const long long per = (ksteps_total + nslices - 1) / nslices;
  const long long ks0 = slice * per;
  const long long ks1 = min(ksteps_total, ks0 + per);
  const int nk = (int)max(0LL, ks1 - ks0);

  extern __shared__ unsigned char g_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(g_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)GSTAGES * G_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GSTAGES;
  uint64_t* acc_bar = empty_bar + GSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmap);
    for (int s = 0; s < GSTAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, GN);   // 256 columns x 128 lanes fp32
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int k = 0; k < nk; k++) {
        const int st = k % GSTAGES;
        if (k >= GSTAGES) mbar_wait(&empty_bar[st], (uint32_t)(((k / GSTAGES) - 1) & 1));
        unsigned char* a_dst = smem + (size_t)st * G_STAGE_BYTES;
        unsigned char* b_dst = a_dst + G_A_BYTES;
        mbar_arrive_expect_tx(&full_bar[st], G_STAGE_BYTES);
        const int krow = (int)((ks0 + k) * GK);
#pragma unroll
        for (int b = 0; b < GM / 64; b++) tma_load_2d(a_dst + b * G_BOX_BYTES, tmap, tile.bi * GM + b * 64, krow, &full_bar[st]);
#pragma unroll
        for (int b = 0; b < GN / 64; b++) tma_load_2d(b_dst + b * G_BOX_BYTES, tmap, tile.bj * GN + b * 64, krow, &full_bar[st]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc_bf16_mn(GM, GN);
      for (int k = 0; k < nk; k++) {
        const int st = k % GSTAGES;
        mbar_wait(&full_bar[st], (uint32_t)((k / GSTAGES) & 1));
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)st * G_STAGE_BYTES);
        const uint32_t b_addr = a_addr + G_A_BYTES;
#pragma unroll
        for (int kk = 0; kk < GK / UK; kk++) {
          // 16 k-rows = 2 swizzle atoms of 1024 B along K
          const uint64_t da = umma_desc_mn_sw128(a_addr + kk * (UK * 128), G_BOX_BYTES, 1024);
          const uint64_t db = umma_desc_mn_sw128(b_addr + kk * (UK * 128), G_BOX_BYTES, 1024);
          umma_f16(tmem_base, da, db, idesc, (k | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[st]);   // frees the smem stage when these MMAs retire
      }
      umma_commit(acc_bar);            // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 -> TMEM lane quadrant (warp % 4) =====
    const int quad = warp & 3;
    float* out = pb.Hpart + (size_t)slice * Dp * Dp;
    const int row = tile.bi * GM + quad * 32 + lane;
    if (nk > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < GN; c0 += 32) {
      uint32_t r[32];
      if (nk > 0) {
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) r[j] = 0u;
      }
      const int col = tile.bj * GN + c0;
      if (row < Dp && col < Dp) {
        float4* dst = reinterpret_cast<float4*>(out + (size_t)row * Dp + col);
#pragma unroll
        for (int j = 0; j < 8; j++)
          dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, GN);
  }
}

// ------------------------------------------------------------------------------------------
// fp32 SIMT debug kernel: same operand (bf16 Xt), same output format (slice 0; other slices zeroed).
// 64x64 lower tiles, 256 threads x (4x4).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_simt_kernel(const Problem* __restrict__ probs, int force) {
  const Problem& pb = probs[blockIdx.z];
  Ctrl* ctrl = pb.ctrl;
  if (!force && (ctrl->done || !ctrl->need_hess)) return;
  if (blockIdx.x > blockIdx.y) return;
  const int Dp = pb.Dp;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  __shared__ float Ai[16][64 + 1];
  __shared__ float Aj[16][64 + 1];
  const int tid = threadIdx.x;
  const int ti = (tid / 16) * 4, tj = (tid % 16) * 4;
  float acc[4][4] = {};
  for (long long r0 = 0; r0 < pb.n; r0 += 16) {
    for (int e = tid; e < 16 * 64; e += 256) {
      const int r = e / 64, c = e % 64;
      const long long rr = r0 + r;
      Ai[r][c] = rr < pb.n ? __bfloat162float(pb.Xt[(size_t)rr * Dp + i0 + c]) : 0.f;
      Aj[r][c] = rr < pb.n ? __bfloat162float(pb.Xt[(size_t)rr * Dp + j0 + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) {
      float x[4], y[4];
#pragma unroll
      for (int a = 0; a < 4; a++) { x[a] = Ai[r][ti + a]; y[a] = Aj[r][tj + a]; }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = fmaf(x[a], y[b], acc[a][b]);
    }
    __syncthreads();
  }
  for (int s = 0; s < pb.gram_slices; s++) {
    float* out = pb.Hpart + (size_t)s * Dp * Dp;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) out[(size_t)(i0 + ti + a) * Dp + j0 + tj + b] = (s == 0) ? acc[a][b] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor map over Xt [n][Dp] bf16 row-major: dim0 = feature (contiguous), dim1 = row; box 64 x GK, 128B swizzle.
int gram_make_tensor_map(void* out_map_host /*CUtensorMap, 128 B*/, const void* xt, long long n, int Dp) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || !p) return 1;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  cuuint64_t dims[2] = {(cuuint64_t)Dp, (cuuint64_t)n};
  cuuint64_t strides[1] = {(cuuint64_t)Dp * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)GK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_map_host), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(xt), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

// Lower block-triangle tile list for a Dp x Dp output (Dp multiple of 128).
int gram_tile_list(int Dp, short* bi_bj_pairs /*[2*max]*/, int max_tiles) {
  int n = 0;
  const int nbi = Dp / GM, nbj = (Dp + GN - 1) / GN;
  for (int bi = 0; bi < nbi; bi++)
    for (int bj = 0; bj < nbj; bj++)
      if (bj * GN <= bi * GM + GM - 1) {
        if (n >= max_tiles) return -1;
        bi_bj_pairs[2 * n] = (short)bi; bi_bj_pairs[2 * n + 1] = (short)bj; n++;
      }
  return n;
}

cudaError_t gram_launch_tcgen05(const Problem* d_probs, int nprob, const void* d_tmaps, const void* d_tiles, int ntiles,
                                int nslices, int force, cudaStream_t st, int* launches) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gram_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  gram_tcgen05_kernel<<<dim3(ntiles, nslices, nprob), G_THREADS, G_SMEM, st>>>(
      d_probs, reinterpret_cast<const CUtensorMap*>(d_tmaps), reinterpret_cast<const GramTile*>(d_tiles), ntiles, force);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t gram_launch_simt(const Problem* d_probs, int nprob, int Dp, int force, cudaStream_t st, int* launches) {
  const int T = Dp / 64;
  gram_simt_kernel<<<dim3(T, T, nprob), 256, 0, st>>>(d_probs, force);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace mlease
