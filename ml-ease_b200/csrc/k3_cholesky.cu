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
  __shared__ double Ai[TB][NB + 1];
  __shared__ double Aj[TB][NB + 1];
  double* H = pb.Lc;
  const int c0 = k * NB;
  const int tid = threadIdx.x;
  for (int e = tid; e < TB * NB; e += 256) {
    const int i = e / NB, kk = e % NB;
    Ai[i][kk] = (i0 + i < ldh) ? H[(size_t)(i0 + i) * ldh + c0 + kk] : 0.0;
    Aj[i][kk] = (j0 + i < ldh) ? H[(size_t)(j0 + i) * ldh + c0 + kk] : 0.0;
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
    for (int a = 0; a < 4; a++) { x[a] = Ai[ti + a][kk]; y[a] = Aj[tj + a][kk]; }
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
    else { c->hess_valid = 1; c->hess_builds++; c->tot_hess++; }
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
  chol_finish_kernel<<<dim3(nb, nprob), 256, 0, st>>>(d_probs);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace mlease
