// k3_cholesky.cu -- K3: H = sum of split-K Gram partials + diag(q), blocked right-looking Cholesky
// in fp64, batched over problems (blockIdx.y / z).  The triangular solves live in newton.cu.
//
// No direct reference equivalent on the ADMM path (TRON is matrix-free, bw/Tron.java:126-179);
// the only Cholesky in the reference is commons-math's at llf/LibLinear.java:321-325 (posterior
// covariance).  H itself is LogisticRegressionL2.hessian (llf/LogisticRegressionL2.java:258-297).
//
// fp64 on purpose: the Gram comes from bf16 tensor-core products, but the factorisation must not
// break down when cond(H) approaches 1/eps_fp32; 3.3e8 flop at D'=1001 is latency- not
// throughput-bound on B200's fp64 pipe.
#include "kernels.cuh"

namespace mlease {

constexpr int NB = 32;   // panel width
constexpr int TB = 64;   // trailing-update tile

// Hd (lower incl. diagonal) = sum_s Hpart[s] + diag(q); padded rows/cols (>= Dt) = identity.
__global__ void chol_prep_kernel(const Problem* __restrict__ probs) {
  const Problem& pb = probs[blockIdx.z];
  Ctrl* c = pb.ctrl;
  if (c->done || !c->need_hess) return;
  const int ldh = pb.ldh, Dt = pb.Dt, Dp = pb.Dp, S = pb.gram_slices;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ldh || j >= ldh || j > i) return;
  double v;
  if (i < Dt) {
    double s = 0.0;
    const size_t off = (size_t)i * Dp + j;
    for (int t = 0; t < S; t++) s += (double)pb.Hpart[(size_t)t * Dp * Dp + off];
    if (i == j) s += pb.q[i];
    v = s;
  } else {
    v = (i == j) ? 1.0 : 0.0;
  }
  pb.Lc[(size_t)i * ldh + j] = v;
}

// Panel step k: every CTA factorises the NBxNB diagonal block redundantly in shared memory
// (cheap), inverts it, and computes its RB rows of L21 = A21 * L11^-T.  CTA 0 stores L11.
constexpr int RB = 64;
__global__ void __launch_bounds__(256) chol_panel_kernel(const Problem* __restrict__ probs, int k) {
  const Problem& pb = probs[blockIdx.y];
  Ctrl* c = pb.ctrl;
  if (c->done || !c->need_hess) return;
  __shared__ double A[NB][NB + 1];
  __shared__ double Li[NB][NB + 1];
  __shared__ double P[RB][NB + 1];
  __shared__ int s_bad;
  const int ldh = pb.ldh;
  const int c0 = k * NB;
  double* H = pb.Lc;
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e / NB, j = e % NB;
    A[i][j] = (j <= i) ? H[(size_t)(c0 + i) * ldh + c0 + j] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < NB; j++) {
    if (tid == 0) {
      const double d = A[j][j];
      if (!(d > 0.0)) { s_bad = 1; A[j][j] = 1.0; } else A[j][j] = sqrt(d);
    }
    __syncthreads();
    const double dj = A[j][j];
    for (int i = j + 1 + tid; i < NB; i += 256) A[i][j] /= dj;
    __syncthreads();
    for (int e = tid; e < NB * NB; e += 256) {
      const int i = e / NB, kk = e % NB;
      if (kk > j && i >= kk) A[i][kk] -= A[i][j] * A[kk][j];
    }
    __syncthreads();
  }
  // inverse of the lower-triangular A: thread cidx solves column cidx
  if (tid < NB) {
    const int cc = tid;
    for (int i = 0; i < cc; i++) Li[i][cc] = 0.0;
    Li[cc][cc] = 1.0 / A[cc][cc];
    for (int i = cc + 1; i < NB; i++) {
      double s = 0.0;
      for (int kk = cc; kk < i; kk++) s += A[i][kk] * Li[kk][cc];
      Li[i][cc] = -s / A[i][i];
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    // The factorised diagonal block goes to a side buffer: sibling CTAs of this launch may still be
    // loading the unfactorised block from H.  chol_finish_kernel copies the blocks back.
    for (int e = tid; e < NB * NB; e += 256) {
      const int i = e / NB, j = e % NB;
      pb.Ldiag[(size_t)(c0 + i) * NB + j] = (j <= i) ? A[i][j] : 0.0;
      pb.Ldinv[(size_t)(c0 + i) * NB + j] = (j <= i) ? Li[i][j] : 0.0;
    }
    if (tid == 0 && s_bad) { c->fail = 1; }
  }
  // rows of the panel below the diagonal block handled by this CTA
  const int r0 = c0 + NB + blockIdx.x * RB;
  if (r0 >= ldh) return;
  const int rows = min(RB, ldh - r0);
  for (int e = tid; e < rows * NB; e += 256) {
    const int i = e / NB, j = e % NB;
    P[i][j] = H[(size_t)(r0 + i) * ldh + c0 + j];
  }
  __syncthreads();
  for (int e = tid; e < rows * NB; e += 256) {
    const int i = e / NB, j = e % NB;
    double s = 0.0;
    for (int kk = 0; kk <= j; kk++) s += P[i][kk] * Li[j][kk];   // (A21 * L11^-T)[i][j]
    H[(size_t)(r0 + i) * ldh + c0 + j] = s;
  }
}

// Trailing update A22 -= L21 L21^T on lower-triangular TBxTB tiles.
__global__ void __launch_bounds__(256) chol_update_kernel(const Problem* __restrict__ probs, int k) {
  const Problem& pb = probs[blockIdx.z];
  Ctrl* c = pb.ctrl;
  if (c->done || !c->need_hess) return;
  if (blockIdx.x > blockIdx.y) return;  // lower tiles only (x = tile col, y = tile row)
  const int ldh = pb.ldh;
  const int base = (k + 1) * NB;
  const int i0 = base + blockIdx.y * TB, j0 = base + blockIdx.x * TB;
  if (i0 >= ldh || j0 >= ldh) return;
  __shared__ double Ai[NB][TB + 2];   // [k][row]: a thread's 4 rows are contiguous -> conflict-light vector reads
  __shared__ double Aj[NB][TB + 2];
  double* H = pb.Lc;
  const int c0 = k * NB;
  const int tid = threadIdx.x;
  for (int e = tid; e < TB * NB; e += 256) {
    const int i = e / NB, kk = e % NB;
    Ai[kk][i] = (i0 + i < ldh) ? H[(size_t)(i0 + i) * ldh + c0 + kk] : 0.0;
    Aj[kk][i] = (j0 + i < ldh) ? H[(size_t)(j0 + i) * ldh + c0 + kk] : 0.0;
  }
  __syncthreads();
  const int ti = (tid / 16) * 4, tj = (tid % 16) * 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < NB; kk++) {
    double x[4], y[4];
#pragma unroll
    for (int a = 0; a < 4; a++) { x[a] = Ai[kk][ti + a]; y[a] = Aj[kk][tj + a]; }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] += x[a] * y[b];
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int i = i0 + ti + a, j = j0 + tj + b;
      if (i < ldh && j <= i) H[(size_t)i * ldh + j] -= acc[a][b];
    }
}

__global__ void chol_finish_kernel(const Problem* __restrict__ probs) {
  const Problem& pb = probs[blockIdx.y];
  Ctrl* c = pb.ctrl;
  if (c->done || !c->need_hess) return;
  const int c0 = blockIdx.x * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    if (j <= i) pb.Lc[(size_t)(c0 + i) * pb.ldh + c0 + j] = pb.Ldiag[(size_t)(c0 + i) * NB + j];
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (c->fail == 1) { c->done = 1; c->hess_valid = 0; }
    else { c->hess_valid = 1; c->hess_builds++; c->tot_hess++; c->bfgs_count = 0; }
  }
}

// ------------------------------------------------------------------------------------------
// Explicit inverse, built once per factorisation so that every chord-Newton direction afterwards is a
// single multi-CTA GEMV instead of two latency-bound triangular solves:
//   Y = L^-1    : forward substitution with 64 right-hand sides (columns of I) per CTA, tiles staged through shared
//                 memory, diagonal blocks applied through their stored inverses
//   Hinv = Y^T Y: 64x64 tiles, both triangles written
// ------------------------------------------------------------------------------------------
// Y = L^-1 by blocked forward substitution, one CTA per NR columns of Y.  Row block kb of those columns is
//   Y[kb] = Ldinv[kb] * ( I[kb] - sum_{jb<kb} L[kb][jb] * Y[jb] )
// with 32x32 tiles of L and 32xNR tiles of the already computed Y (read back from global memory / L2, so the column
// count per CTA is not bounded by shared memory: D' = 10k works the same way as D' = 1k).  The next pair of tiles is
// prefetched into registers while the current pair is multiplied.  Thread (ti, tq) owns row ti and the NR/8 columns
// tq, tq+8, tq+16, ... : for a fixed k the 8 lanes of a row read 8 consecutive doubles (no bank conflicts).
template <int NR>
__global__ void __launch_bounds__(256) trinv_kernel(const Problem* __restrict__ probs) {
  const Problem& pb = probs[blockIdx.y];
  Ctrl* c = pb.ctrl;
  if (c->done || !c->need_hess) return;
  constexpr int CPT = NR / 8;   // columns per thread
  __shared__ double Lt[NB][NB + 1];
  __shared__ double Yt[NB][NR + 2];
  __shared__ double Rb[NB][NR + 2];
  const int ldh = pb.ldh, nb = ldh / NB;
  const int c0 = blockIdx.x * NR;
  if (c0 >= ldh) return;
  const int ncols = min(NR, ldh - c0);
  const int kb0 = c0 / NB;
  const int tid = threadIdx.x;
  const int ti = tid >> 3, tq = tid & 7;
  const double* __restrict__ L = pb.Lc;
  double* __restrict__ Y = pb.Yinv;
  const int lr = tid >> 3, lc = (tid & 7) * 4;   // Lt 32x32: 4 consecutive doubles per thread
  for (int kb = kb0; kb < nb; kb++) {
    const int r0 = kb * NB;
    double acc[CPT];
#pragma unroll
    for (int q = 0; q < CPT; q++) acc[q] = (r0 + ti == c0 + tq + 8 * q) ? 1.0 : 0.0;
    double pl[4], py[CPT];
    auto prefetch = [&](int jb) {
#pragma unroll
      for (int q = 0; q < 4; q++) pl[q] = L[(size_t)(r0 + lr) * ldh + jb * NB + lc + q];
#pragma unroll
      for (int q = 0; q < CPT; q++) py[q] = (tq + 8 * q < ncols) ? Y[(size_t)(jb * NB + ti) * ldh + c0 + tq + 8 * q] : 0.0;
    };
    if (kb > kb0) prefetch(kb0);
    for (int jb = kb0; jb < kb; jb++) {
#pragma unroll
      for (int q = 0; q < 4; q++) Lt[lr][lc + q] = pl[q];
#pragma unroll
      for (int q = 0; q < CPT; q++) Yt[ti][tq + 8 * q] = py[q];
      __syncthreads();
      if (jb + 1 < kb) prefetch(jb + 1);
#pragma unroll 8
      for (int kk = 0; kk < NB; kk++) {
        const double l = Lt[ti][kk];
#pragma unroll
        for (int q = 0; q < CPT; q++) acc[q] -= l * Yt[kk][tq + 8 * q];
      }
      __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < CPT; q++) Rb[ti][tq + 8 * q] = acc[q];
#pragma unroll
    for (int q = 0; q < 4; q++) Lt[lr][lc + q] = pb.Ldinv[(size_t)(r0 + lr) * NB + lc + q];
    __syncthreads();
    double yv[CPT];
#pragma unroll
    for (int q = 0; q < CPT; q++) yv[q] = 0.0;
    for (int kk = 0; kk <= ti; kk++) {
      const double l = Lt[ti][kk];
#pragma unroll
      for (int q = 0; q < CPT; q++) yv[q] += l * Rb[kk][tq + 8 * q];
    }
#pragma unroll
    for (int q = 0; q < CPT; q++)
      if (tq + 8 * q < ncols) Y[(size_t)(r0 + ti) * ldh + c0 + tq + 8 * q] = yv[q];
    __syncthreads();   // the Y block just written is read back (by other threads of this CTA) for the next row blocks
  }
}

__global__ void __launch_bounds__(256) hinv_syrk_kernel(const Problem* __restrict__ probs) {
  const Problem& pb = probs[blockIdx.z];
  Ctrl* c = pb.ctrl;
  if (c->done || !c->need_hess) return;
  if (blockIdx.x > blockIdx.y) return;
  const int ldh = pb.ldh;
  const int i0 = blockIdx.y * TB, j0 = blockIdx.x * TB;   // i0 >= j0
  if (i0 >= ldh) return;
  __shared__ double Yi[NB][TB + 1];
  __shared__ double Yj[NB][TB + 1];
  const double* Y = pb.Yinv;
  const int tid = threadIdx.x;
  const int ti = (tid / 16) * 4, tj = (tid % 16) * 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
  // Y is lower triangular: Y[k][i] != 0 only for k >= i, so start at the row block of i0 (>= j0)
  for (int k0 = (i0 / NB) * NB; k0 < ldh; k0 += NB) {
    for (int e = tid; e < NB * TB; e += 256) {
      const int kk = e / TB, cc = e % TB;
      const int k = k0 + kk;
      Yi[kk][cc] = (i0 + cc < ldh && i0 + cc <= k) ? Y[(size_t)k * ldh + i0 + cc] : 0.0;
      Yj[kk][cc] = (j0 + cc < ldh && j0 + cc <= k) ? Y[(size_t)k * ldh + j0 + cc] : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < NB; kk++) {
      double x[4], y[4];
#pragma unroll
      for (int a = 0; a < 4; a++) { x[a] = Yi[kk][ti + a]; y[a] = Yj[kk][tj + a]; }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] += x[a] * y[b];
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int i = i0 + ti + a, j = j0 + tj + b;
      if (i < ldh && j < ldh) {
        pb.Hinv[(size_t)i * ldh + j] = acc[a][b];
        pb.Hinv[(size_t)j * ldh + i] = acc[a][b];
      }
    }
}

cudaError_t cholesky_launch(const Problem* d_probs, int nprob, int ldh, cudaStream_t st, int* launches) {
  {
    dim3 blk(32, 8);
    dim3 grd((ldh + 31) / 32, (ldh + 7) / 8, nprob);
    chol_prep_kernel<<<grd, blk, 0, st>>>(d_probs);
    if (launches) *launches += 1;
  }
  const int nb = ldh / NB;
  for (int k = 0; k < nb; k++) {
    const int below = ldh - (k + 1) * NB;
    const int gx = below > 0 ? (below + RB - 1) / RB : 1;
    chol_panel_kernel<<<dim3(gx, nprob), 256, 0, st>>>(d_probs, k);
    if (launches) *launches += 1;
    if (below > 0) {
      const int T = (below + TB - 1) / TB;
      chol_update_kernel<<<dim3(T, T, nprob), 256, 0, st>>>(d_probs, k);
      if (launches) *launches += 1;
    }
  }
  // explicit inverse (reads the panel blocks below the diagonal from Lc and the diagonal inverses from Ldinv)
  {
    if (ldh <= 2048) trinv_kernel<32><<<dim3((ldh + 31) / 32, nprob), 256, 0, st>>>(d_probs);   // more CTAs for small systems
    else trinv_kernel<64><<<dim3((ldh + 63) / 64, nprob), 256, 0, st>>>(d_probs);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (launches) *launches += 1;
    const int T = (ldh + TB - 1) / TB;
    hinv_syrk_kernel<<<dim3(T, T, nprob), 256, 0, st>>>(d_probs);
    if (launches) *launches += 1;
  }
  chol_finish_kernel<<<dim3(nb, nprob), 256, 0, st>>>(d_probs);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace mlease
