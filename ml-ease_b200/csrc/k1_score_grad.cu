// k1_score_grad.cu -- K1: fused score + sigmoid/IRLS reweight + gradient (+ objective, + the
// sqrt(d)-scaled bf16 operand of the Gram kernel) in ONE pass over X.
//
// Replaces LogisticRegressionL2.fun + grad (llf/LogisticRegressionL2.java:156-225), i.e. the
// two sparse passes Xv (:115-129) and XTv (:131-150), and the score recomputation inside
// hessian() (:261-269).  HBM-bound: algorithmic bytes per row = 4*ldx (X once) + 9 (y,w,o).
//
// Dense layout: X row-major [n][ldx] fp32, ldx % 4 == 0, bias column physical.  A CTA streams
// row tiles of R rows (R*ldx*4 contiguous bytes) through an S-stage shared-memory ring with
// 1-D bulk TMA copies (cp.async.bulk, mbarrier complete_tx), then
//   phase A: one warp per row, float4 smem reads, s_i = x_i.beta + o_i, p = sigmoid(y s),
//            r_i = w (p-1) y, d_i = w p (1-p), loss_i                      (row-dot)
//   phase B: one thread per float4 column group, g += r_i * x_i  over the tile's rows, and
//            (optionally) Xt[i][:] = bf16(sqrt(d_i) * x_i[:])            (column-sum + emit)
// so every element of X is read from HBM exactly once and from shared memory twice.
// Per-CTA partial gradients are accumulated in fp64 registers and written to gpart; a fixed
// order reduction (k1_reduce_decide in newton.cu) makes the result run-to-run deterministic.
#include "kernels.cuh"

namespace mlease {

constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;

template <int G>
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_dense_kernel(const Problem* __restrict__ probs, int R, int S, int force_emit) {
  const Problem& pb = probs[blockIdx.y];
  Ctrl* ctrl = pb.ctrl;
  if (ctrl->done) return;
  const bool emit = force_emit >= 0 ? (force_emit != 0) : (ctrl->emit != 0);

  const int ldx = pb.ldx;
  const int ncg = ldx >> 2;  // float4 column groups
  const long long n = pb.n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const size_t stage_bytes = (size_t)R * ldx * sizeof(float);
  float* stage0 = reinterpret_cast<float*>(smem_raw);
  float* beta_s = reinterpret_cast<float*>(smem_raw + (size_t)S * stage_bytes);
  float* r_s = beta_s + ldx;
  float* sd_s = r_s + R;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sd_s + R) + 15) & ~uintptr_t(15));
  double* red_s = reinterpret_cast<double*>(full_bar + 8);  // 8 warps of scratch

  const long long ntiles = (n + R - 1) / R;
  // tiles handled by this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const long long my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < S; s++) mbar_init(&full_bar[s], 1);
    fence_mbar_init();
  }
  for (int c = tid; c < ldx; c += K1_THREADS) beta_s[c] = pb.beta_tf[c];
  __syncthreads();

  auto issue = [&](long long k) {  // tile index within this CTA's sequence -> stage k % S
    const long long t = blockIdx.x + k * (long long)gridDim.x;
    const long long row0 = t * R;
    const int rows = (int)min((long long)R, n - row0);
    const uint32_t bytes = (uint32_t)((size_t)rows * ldx * sizeof(float));
    uint64_t* bar = &full_bar[k % S];
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(reinterpret_cast<unsigned char*>(stage0) + (k % S) * stage_bytes, pb.X + row0 * ldx, bytes, bar);
  };
  if (tid == 0) {
    for (int k = 0; k < S - 1 && k < my_tiles; k++) issue(k);
  }

  // thread -> column-group mapping for phase B
  int nsl = 1, sl = 0, cg0 = tid;
  bool activeB = true;
  if (G == 1) {
    nsl = K1_THREADS / ncg;
    if (nsl < 1) nsl = 1;
    sl = tid / ncg;
    cg0 = tid - sl * ncg;
    activeB = sl < nsl;
  }
  double acc64[G][4];
#pragma unroll
  for (int g = 0; g < G; g++) acc64[g][0] = acc64[g][1] = acc64[g][2] = acc64[g][3] = 0.0;
  double loss64 = 0.0;

  const float4* beta4 = reinterpret_cast<const float4*>(beta_s);

  for (long long k = 0; k < my_tiles; k++) {
    const int st = (int)(k % S);
    if (tid == 0 && k + S - 1 < my_tiles) issue(k + S - 1);
    mbar_wait(&full_bar[st], (uint32_t)((k / S) & 1));
    const float* tile = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(stage0) + st * stage_bytes);
    const long long row0 = (blockIdx.x + k * (long long)gridDim.x) * R;
    const int rows = (int)min((long long)R, n - row0);

    // ---- phase A: row dots, two rows per warp per step (beta float4 shared by both rows) ---------------
    // lanes 0/1 own the per-row scalars of rows r/r+1: their y,w,o loads are issued BEFORE the dot loop so the
    // global-load latency hides behind it; the sigmoid/loss math then runs on two lanes at once.
    for (int r = 2 * warp; r < rows; r += 2 * K1_WARPS) {
      const bool two = (r + 1) < rows;
      const int myr = r + (lane & 1);
      float yy = 0.f, ww = 0.f, oo = 0.f;
      if (lane < 2 && myr < rows) {
        const long long i = row0 + myr;
        yy = (float)pb.y[i]; ww = pb.w[i]; oo = pb.o[i];
      }
      const float4* x0p = reinterpret_cast<const float4*>(tile + (size_t)r * ldx);
      const float4* x1p = reinterpret_cast<const float4*>(tile + (size_t)(two ? r + 1 : r) * ldx);
      float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
      int c = lane;
      for (; c + 32 < ncg; c += 64) {
        const float4 b0 = beta4[c], b1 = beta4[c + 32];
        const float4 x0 = x0p[c], x1 = x1p[c], x2 = x0p[c + 32], x3 = x1p[c + 32];
        a0 = fmaf(x0.x, b0.x, a0); a0 = fmaf(x0.y, b0.y, a0); a0 = fmaf(x0.z, b0.z, a0); a0 = fmaf(x0.w, b0.w, a0);
        a1 = fmaf(x1.x, b0.x, a1); a1 = fmaf(x1.y, b0.y, a1); a1 = fmaf(x1.z, b0.z, a1); a1 = fmaf(x1.w, b0.w, a1);
        c0 = fmaf(x2.x, b1.x, c0); c0 = fmaf(x2.y, b1.y, c0); c0 = fmaf(x2.z, b1.z, c0); c0 = fmaf(x2.w, b1.w, c0);
        c1 = fmaf(x3.x, b1.x, c1); c1 = fmaf(x3.y, b1.y, c1); c1 = fmaf(x3.z, b1.z, c1); c1 = fmaf(x3.w, b1.w, c1);
      }
      if (c < ncg) {
        const float4 b0 = beta4[c];
        const float4 x0 = x0p[c], x1 = x1p[c];
        a0 = fmaf(x0.x, b0.x, a0); a0 = fmaf(x0.y, b0.y, a0); a0 = fmaf(x0.z, b0.z, a0); a0 = fmaf(x0.w, b0.w, a0);
        a1 = fmaf(x1.x, b0.x, a1); a1 = fmaf(x1.y, b0.y, a1); a1 = fmaf(x1.z, b0.z, a1); a1 = fmaf(x1.w, b0.w, a1);
      }
      const float s0 = warp_sum(a0 + c0), s1 = warp_sum(a1 + c1);
      if (lane < 2 && myr < rows) {
        const float t = yy * ((lane ? s1 : s0) + oo);
        const float e = __expf(-fabsf(t));                 // in (0,1]; ex2.approx path
        const float inv = __frcp_rn(1.f + e);
        const float p = t >= 0.f ? inv : e * inv;          // sigmoid(y s)
        const float qq = t >= 0.f ? e * inv : inv;         // 1 - p, no cancellation
        // log1p(e) = -log(1/(1+e)); absolute error ~1e-7 per row is far below the objective's use (line search only)
        loss64 += (double)(ww * ((t >= 0.f ? 0.f : -t) - __logf(inv)));
        r_s[myr] = -ww * yy * qq;                          // w (p-1) y
        sd_s[myr] = sqrtf(ww * p * qq);                    // sqrt(d_i)
      }
    }
    __syncthreads();

    // ---- phase B: column sums (+ emit scaled bf16 copy) ------------------------------------
    if (activeB) {
#pragma unroll
      for (int g = 0; g < G; g++) {
        const int cg = cg0 + g * K1_THREADS;
        if (cg < ncg) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4* t4 = reinterpret_cast<const float4*>(tile) + cg;
          if (emit) {
            __nv_bfloat16* xt = pb.Xt + (size_t)row0 * pb.Dp + 4 * cg;
            for (int r = sl; r < rows; r += nsl) {
              const float4 x = t4[(size_t)r * ncg];
              const float rr = r_s[r], sd = sd_s[r];
              a.x = fmaf(x.x, rr, a.x); a.y = fmaf(x.y, rr, a.y); a.z = fmaf(x.z, rr, a.z); a.w = fmaf(x.w, rr, a.w);
              __nv_bfloat162 lo = __floats2bfloat162_rn(x.x * sd, x.y * sd);
              __nv_bfloat162 hi = __floats2bfloat162_rn(x.z * sd, x.w * sd);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&lo);
              pk.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(xt + (size_t)r * pb.Dp) = pk;
            }
          } else {
            float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1, a3 = a1;
            int r = sl;
            for (; r + 3 * nsl < rows; r += 4 * nsl) {
              const float4 x0 = t4[(size_t)r * ncg], x1 = t4[(size_t)(r + nsl) * ncg], x2 = t4[(size_t)(r + 2 * nsl) * ncg],
                           x3 = t4[(size_t)(r + 3 * nsl) * ncg];
              const float r0 = r_s[r], r1 = r_s[r + nsl], r2 = r_s[r + 2 * nsl], r3 = r_s[r + 3 * nsl];
              a.x = fmaf(x0.x, r0, a.x); a.y = fmaf(x0.y, r0, a.y); a.z = fmaf(x0.z, r0, a.z); a.w = fmaf(x0.w, r0, a.w);
              a1.x = fmaf(x1.x, r1, a1.x); a1.y = fmaf(x1.y, r1, a1.y); a1.z = fmaf(x1.z, r1, a1.z); a1.w = fmaf(x1.w, r1, a1.w);
              a2.x = fmaf(x2.x, r2, a2.x); a2.y = fmaf(x2.y, r2, a2.y); a2.z = fmaf(x2.z, r2, a2.z); a2.w = fmaf(x2.w, r2, a2.w);
              a3.x = fmaf(x3.x, r3, a3.x); a3.y = fmaf(x3.y, r3, a3.y); a3.z = fmaf(x3.z, r3, a3.z); a3.w = fmaf(x3.w, r3, a3.w);
            }
            for (; r < rows; r += nsl) {
              const float4 x = t4[(size_t)r * ncg];
              const float rr = r_s[r];
              a.x = fmaf(x.x, rr, a.x); a.y = fmaf(x.y, rr, a.y); a.z = fmaf(x.z, rr, a.z); a.w = fmaf(x.w, rr, a.w);
            }
            a.x += (a1.x + a2.x) + a3.x; a.y += (a1.y + a2.y) + a3.y; a.z += (a1.z + a2.z) + a3.z; a.w += (a1.w + a2.w) + a3.w;
          }
          acc64[g][0] += (double)a.x; acc64[g][1] += (double)a.y; acc64[g][2] += (double)a.z; acc64[g][3] += (double)a.w;
        }
      }
    }
    __syncthreads();  // all reads of this stage (and of r_s/sd_s) done -> stage may be refilled
  }

  // ---- CTA epilogue: reduce slices, write partials -------------------------------------------
  double* gp = pb.gpart + (size_t)blockIdx.x * ldx;
  if (G == 1 && nsl > 1) {
    double* sc = reinterpret_cast<double*>(smem_raw);  // all bulk copies have completed and been consumed
    if (activeB) {
      double* d = sc + ((size_t)sl * ncg + cg0) * 4;
      d[0] = acc64[0][0]; d[1] = acc64[0][1]; d[2] = acc64[0][2]; d[3] = acc64[0][3];
    }
    __syncthreads();
    for (int c = tid; c < ldx; c += K1_THREADS) {
      double s = 0.0;
      for (int q = 0; q < nsl; q++) s += sc[(size_t)q * ldx + c];
      gp[c] = s;
    }
  } else {
#pragma unroll
    for (int g = 0; g < G; g++) {
      const int cg = cg0 + g * K1_THREADS;
      if (activeB && cg < ncg) {
        gp[4 * cg + 0] = acc64[g][0]; gp[4 * cg + 1] = acc64[g][1];
        gp[4 * cg + 2] = acc64[g][2]; gp[4 * cg + 3] = acc64[g][3];
      }
    }
  }
  loss64 += __shfl_down_sync(0xffffffffu, loss64, 1);   // lanes 0 and 1 carry the per-row losses
  if (lane == 0) red_s[warp] = loss64;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int wq = 0; wq < K1_WARPS; wq++) s += red_s[wq];
    pb.fpart[blockIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------
// CSR variant (configs 3/4): one warp per row, gather for the score, fp64 atomics for the
// scatter into gpart[0][:] (zeroed by k1_csr_zero).  The bias is implicit (value 1, column Dt-1).
// If emit: the dense scaled row is assembled into Xt (zero-filled by the caller once; each
// refresh rewrites exactly the row's nnz + bias positions).
// ------------------------------------------------------------------------------------------
__global__ void k1_csr_zero_kernel(const Problem* __restrict__ probs) {
  const Problem& pb = probs[blockIdx.y];
  if (pb.ctrl->done) return;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < pb.ldx; c += gridDim.x * blockDim.x) pb.gpart[c] = 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) pb.fpart[0] = 0.0;
}

__global__ void __launch_bounds__(256)
k1_csr_kernel(const Problem* __restrict__ probs, int has_bias, int force_emit) {
  const Problem& pb = probs[blockIdx.y];
  Ctrl* ctrl = pb.ctrl;
  if (ctrl->done) return;
  const bool emit = force_emit >= 0 ? (force_emit != 0) : (ctrl->emit != 0);
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float* __restrict__ bt = pb.beta_tf;
  const int Dt = pb.Dt;
  double loss = 0.0;
  for (long long i = warp_global; i < pb.n; i += nwarps) {
    const long long j0 = pb.rowptr[i], j1 = pb.rowptr[i + 1];
    float a = 0.f;
    for (long long j = j0 + lane; j < j1; j += 32) a = fmaf(pb.vals[j], bt[pb.colidx[j]], a);
    a = warp_sum(a);
    if (has_bias) a += bt[Dt - 1];
    const float yy = (float)pb.y[i], ww = pb.w[i];
    const float t = yy * (a + pb.o[i]);
    const float e = expf(-fabsf(t));
    const float inv = 1.f / (1.f + e);
    const float p = t >= 0.f ? inv : e * inv;
    const float qq = t >= 0.f ? e * inv : inv;
    if (lane == 0) loss += (double)(ww * ((t >= 0.f ? 0.f : -t) + log1pf(e)));
    const float rr = -ww * yy * qq;
    const float sd = sqrtf(ww * p * qq);
    for (long long j = j0 + lane; j < j1; j += 32) {
      const int c = pb.colidx[j];
      const float v = pb.vals[j];
      atomicAdd(&pb.gpart[c], (double)(v * rr));
    }
    if (has_bias && lane == 0) atomicAdd(&pb.gpart[Dt - 1], (double)rr);
    if (emit) {
      __nv_bfloat16* xt = pb.Xt + (size_t)i * pb.Dp;
      // duplicates within a row add in the reference's Xv; the dense assembly must add too
      for (long long j = j0 + lane; j < j1; j += 32) xt[pb.colidx[j]] = __float2bfloat16_rn(0.f);
      __syncwarp();
      for (long long j = j0; j < j1; j++) {  // serial over the row's nnz keeps duplicate handling exact
        if (lane == 0) {
          const int c = pb.colidx[j];
          xt[c] = __float2bfloat16_rn(__bfloat162float(xt[c]) + pb.vals[j] * sd);
        }
      }
      if (has_bias && lane == 0) xt[Dt - 1] = __float2bfloat16_rn(sd);
    }
  }
  if (lane == 0 && loss != 0.0) atomicAdd(&pb.fpart[0], loss);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static size_t k1_smem_bytes(int ldx, int R, int S) {
  return (size_t)S * R * ldx * 4 + (size_t)ldx * 4 + (size_t)2 * R * 4 + 16 + 8 * 8 + 8 * 8 + 64;
}

// Picks the tile height R and stage count S for a given ldx (device smem budget 227 KB).
bool k1_dense_plan(int ldx, int* R_out, int* S_out, int* G_out, size_t* smem_out) {
  const size_t budget = 220 * 1024;
  int G = (ldx / 4 + K1_THREADS - 1) / K1_THREADS;
  if (G > 4) return false;
  if (G == 3) G = 4;
  int S = 3;
  int R = (int)((budget - (size_t)ldx * 4 - 1024) / ((size_t)S * ldx * 4));
  if (R > 64) R = 64;
  if (R >= 8) R &= ~7;
  if (R < 2) return false;
  *R_out = R; *S_out = S; *G_out = G; *smem_out = k1_smem_bytes(ldx, R, S);
  return true;
}

cudaError_t k1_launch(const Problem* d_probs, int nprob, bool csr, int ldx, int has_bias, int ctas_per_problem,
                      int force_emit, cudaStream_t stream, int* launches) {
  if (csr) {
    k1_csr_zero_kernel<<<dim3(4, nprob), 256, 0, stream>>>(d_probs);
    k1_csr_kernel<<<dim3(ctas_per_problem, nprob), 256, 0, stream>>>(d_probs, has_bias, force_emit);
    if (launches) *launches += 2;
    return cudaGetLastError();
  }
  int R, S, G;
  size_t smem;
  if (!k1_dense_plan(ldx, &R, &S, &G, &smem)) return cudaErrorInvalidValue;
  dim3 grid(ctas_per_problem, nprob);
  cudaError_t e;
#define K1_LAUNCH(GG)                                                                                         \
  e = cudaFuncSetAttribute(k1_dense_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
  if (e != cudaSuccess) return e;                                                                              \
  k1_dense_kernel<GG><<<grid, K1_THREADS, smem, stream>>>(d_probs, R, S, force_emit);
  if (G == 1) { K1_LAUNCH(1) } else if (G == 2) { K1_LAUNCH(2) } else { K1_LAUNCH(4) }
#undef K1_LAUNCH
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace mlease
