// aux_kernels.cuh -- non-template kernels (compiled once, in yalps_b200.cu): standalone node assembly (K3),
// input generators (K5), density probe, roundToPrecision probe, shared-memory stream benchmark.
#pragma once

#include "fastdiv.cuh"
#include "kernels.cuh"

namespace yalps {

// K3 standalone: applyCuts (src/branchAndCut.ts:22-61) for a wave of nodes into HBM working copies
// (row stride W, node stride Hcap*W) -- used in front of the grid kernel; K1/K2 fuse the same assembly.
__global__ void k_assemble_nodes(long long n, int rootH, int W, int Hcap, const double *root, const int *root_pos,
                                 const int *cut_off, const double *cut_sign, const int *cut_var, const double *cut_val,
                                 double *work) {
  const size_t root_cells = (size_t)rootH * W;
  for (long long node = blockIdx.y; node < n; node += gridDim.y) {
    double *dst = work + (size_t)node * Hcap * W;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < root_cells; k += (size_t)gridDim.x * blockDim.x)
      dst[k] = root[k];
    const int cbeg = cut_off[node], ncuts = cut_off[node + 1] - cbeg;
    for (int i = blockIdx.x; i < ncuts; i += gridDim.x) {
      const double sign = cut_sign[cbeg + i], value = cut_val[cbeg + i];
      const int p = root_pos[cut_var[cbeg + i]];
      double *dr = dst + (size_t)(rootH + i) * W;
      if (p < W) {
        for (int c = threadIdx.x; c < W; c += blockDim.x) dr[c] = (c == 0) ? __dmul_rn(sign, value) : (c == p ? sign : 0.0);
      } else {
        const double *sr = root + (size_t)(p - W) * W;
        for (int c = threadIdx.x; c < W; c += blockDim.x)
          dr[c] = (c == 0) ? __dmul_rn(sign, __dsub_rn(value, sr[0])) : __dmul_rn(-sign, sr[c]);
      }
    }
  }
}

// ---- K5: generators -------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t prospector_hash(uint32_t x) {  // tests/helpers/util.ts:20-29
  x ^= x >> 16;
  x *= 0x21f0aaadu;
  x ^= x >> 15;
  x *= 0xd35a2d97u;
  x ^= x >> 15;
  return x;
}

// draw d (0-based) of newRand(seed0): state after d+1 increments (tests/helpers/util.ts:38-41)
__host__ __device__ __forceinline__ double rand_draw(uint32_t seed0, uint32_t d) {
  return (double)prospector_hash(seed0 + (d + 1u) * 0x9e3779b9u) / 4294967296.0;
}

// Dense synthetic LPs.  Draw order: c_1..c_n, then per row a_k1..a_kn, b_k.
__global__ void k_generate_synthetic(long long first, long long n, int m, int nvars, int neg_rows, uint32_t salt,
                                     double *out) {
  const int W = nvars + 1, H = m + 1;
  const size_t cells = (size_t)W * H;
  const size_t total = (size_t)n * cells;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const long long i = (long long)(g / cells);
    const int cell = (int)(g % cells);
    const int r = cell / W, c = cell % W;
    const uint32_t seed0 = prospector_hash((uint32_t)(first + i) ^ salt);
    double v;
    if (r == 0) {
      v = (c == 0) ? 0.0 : rand_draw(seed0, (uint32_t)(c - 1));
    } else {
      const uint32_t base = (uint32_t)nvars + (uint32_t)(r - 1) * (uint32_t)(nvars + 1);
      if (c == 0) {
        const double u = rand_draw(seed0, base + (uint32_t)nvars);
        v = (r <= neg_rows) ? -(0.5 + u) : (double)nvars * (0.25 + 0.5 * u);
      } else {
        v = rand_draw(seed0, base + (uint32_t)(c - 1));
        if (r <= neg_rows) v = -v;
      }
    }
    out[g] = v;
  }
}

// RHS-perturbed replicas of one base tableau.
__global__ void k_generate_replicas(long long first, long long n, int H, int W, const double *base, const int *group,
                                    double eps, uint32_t salt, double *out) {
  const size_t cells = (size_t)W * H;
  const size_t total = (size_t)n * cells;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const long long i = (long long)(g / cells);
    const int cell = (int)(g % cells);
    const int r = cell / W, c = cell % W;
    double v = base[cell];
    if (c == 0 && group[r] >= 0) {
      const uint32_t seed0 = prospector_hash((uint32_t)(first + i) ^ salt);
      const double u = rand_draw(seed0, (uint32_t)group[r]);
      v = __dmul_rn(v, __dadd_rn(1.0, __dmul_rn(eps, __dsub_rn(__dmul_rn(2.0, u), 1.0))));
    }
    out[g] = v;
  }
}

// Replicas of one base tableau that differ only in the RHS column (yalps_solve_replicas): the caller ships the base
// once and n*H right-hand sides; the working copies are assembled here, in HBM, at copy bandwidth.
__global__ void k_expand_replicas(long long n, int H, int W, const double *base, const double *rhs, double *out) {
  const size_t cells = (size_t)W * H;
  const size_t total = (size_t)n * cells;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const size_t i = g / cells;
    const int cell = (int)(g - i * cells);
    const int r = cell / W, c = cell - r * W;
    out[g] = c == 0 ? rhs[i * (size_t)H + r] : base[cell];
  }
}

// Incumbent min-allreduce, the part inside one GPU: the logical ranks that share this GPU are reduced by ONE kernel
// over all their slots (never by launches that wait for one another); slot[k] then goes through ncclAllReduce(min)
// across the distinct GPUs and is broadcast back to the rank slots.
__global__ void k_incumbent_reduce(double *slots, int k) {
  double m = slots[0];
  for (int i = 1; i < k; i++) m = slots[i] < m ? slots[i] : m;
  slots[k] = m;
}
__global__ void k_incumbent_broadcast(double *slots, int k) {
  const double m = slots[k];
  for (int i = 0; i < k; i++) slots[i] = m;
}

// Non-zero count of a sample of one tableau (kernel-path policy: sparse batches go to the HBM/L2-resident kernel).
__global__ void k_sample_density(const double *m, long long cells, long long step, int *out /* [2]: seen, nz */) {
  int seen = 0, nz = 0;
  for (long long k = (long long)threadIdx.x * step; k < cells; k += (long long)blockDim.x * step) {
    seen++;
    nz += m[k] != 0.0;
  }
  atomicAdd(out, seen);
  atomicAdd(out + 1, nz);
}

// fastdiv.cuh against __ddiv_rn on hash-generated operand pairs; counts bit mismatches.
//   mode 0: random bit patterns (all exponents, NaN/inf included)   mode 1: tableau-like magnitudes
//   mode 2: special numerators (0, -0, inf, NaN, denormals, huge)     mode 3: exact quotients and small integers
//   mode 4: exponents near the denormal / overflow boundaries
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}

__global__ void k_probe_division(long long n, unsigned long long seed, int mode, unsigned long long *out /* [4] */) {
  unsigned long long bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long a = splitmix64(seed + 2 * (unsigned long long)i), b = splitmix64(seed + 2 * (unsigned long long)i + 1);
    double x = __longlong_as_double((long long)a), y = __longlong_as_double((long long)b);
    if (mode == 1) {
      x = ((double)(a >> 11) / 9007199254740992.0 - 0.5) * (double)(1 + (a & 1023));
      y = ((double)(b >> 11) / 9007199254740992.0 - 0.5) * (double)(1 + (b & 255));
      if ((a & 0x7000) == 0) x = 0.0;
      if ((a & 0x7000) == 0x1000) x = -0.0;
    } else if (mode == 2) {
      const double sp[12] = {0.0, -0.0, d_inf(), -d_inf(), d_nan(), 4.9e-324, -4.9e-324, 2.2250738585072014e-308,
                             1.7976931348623157e308, -1.7976931348623157e308, 1e-16, 1.0};
      x = sp[a % 12];
      if ((b & 15) == 0) y = sp[(b >> 4) % 12];
    } else if (mode == 3) {
      const double k = (double)((long long)(a & 0xffff) - 32768);
      y = (double)((long long)(b & 0xfffff) - 524288) * 0.125;
      x = (a & 0x10000) ? __dmul_rn(k, y) : k;
    } else if (mode == 4) {
      const unsigned long long ex = (a & 1) ? (unsigned long long)(a >> 1 & 63) : (unsigned long long)(0x7fe - (a >> 1 & 63));
      const unsigned long long ey = (b & 1) ? (unsigned long long)(b >> 1 & 63) : (unsigned long long)(0x7fe - (b >> 1 & 63));
      x = __longlong_as_double((long long)((a & 0x800fffffffffffffULL) | (ex << 52)));
      y = __longlong_as_double((long long)((b & 0x800fffffffffffffULL) | (ey << 52)));
    }
    const Recip rc(y);
    const double q1 = rc.quot(x), q2 = rc.quot(y), q3 = div_rn(1.0, y);
    const double r1 = __ddiv_rn(x, y), r2 = __ddiv_rn(y, y), r3 = __ddiv_rn(1.0, y);
    // the branch-free batch form (K1t): accepted quotients must already be exact, rejected ones go out of line
    RecipBatch rb(y);
    double b1 = rb.quot(x), b2 = rb.quot(y), b3 = rb.quot(1.0);
    (void)rb.quot(__longlong_as_double((long long)(a ^ b)), false);  // an unused quotient must not change the verdict
    if (!rb.ok) {
      b1 = div_rn_slow(x, y);
      b2 = div_rn_slow(y, y);
      b3 = div_rn_slow(1.0, y);
    }
    if (__double_as_longlong(q1) != __double_as_longlong(r1) || __double_as_longlong(q2) != __double_as_longlong(r2) ||
        __double_as_longlong(q3) != __double_as_longlong(r3) || __double_as_longlong(b1) != __double_as_longlong(r1) ||
        __double_as_longlong(b2) != __double_as_longlong(r2) || __double_as_longlong(b3) != __double_as_longlong(r3)) {
      bad++;
      if (atomicAdd(out + 1, 1ULL) == 0) {
        out[2] = (unsigned long long)__double_as_longlong(x);
        out[3] = (unsigned long long)__double_as_longlong(y);
      }
    }
  }
  if (bad) atomicAdd(out, bad);
}

__global__ void k_round_to_precision(long long n, const double *x, double precision, double *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = round_to_precision(x[i], precision);
}

// L2 stream: every thread reads and rewrites 16-byte cells of a buffer that fits L2, `iters` times over
// (the ld - mul - sub - st pattern of the rank-1 update on an L2-resident working copy).  bytes = words*8 * iters * 2.
__global__ void k_l2_stream(double2 *buf, size_t n2, int iters, double coef) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; it++)
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) {
      double2 x = __ldcg(buf + k);
      x.x = __dsub_rn(x.x, __dmul_rn(coef, 1e-9));
      x.y = __dsub_rn(x.y, __dmul_rn(coef, 2e-9));
      __stcg(buf + k, x);
    }
}

// Shared-memory stream: every thread reads and rewrites 8-byte cells of a CTA-private buffer.
// bytes = grid * iters * words * 16.
__global__ void k_smem_stream(int words, int iters, double *sink) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *buf = reinterpret_cast<double *>(smem_raw);
  for (int k = threadIdx.x; k < words; k += blockDim.x) buf[k] = (double)k;
  __syncthreads();
  double acc = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll 4
    for (int k = threadIdx.x; k < words; k += blockDim.x) {
      const double x = buf[k];
      buf[k] = x + 1.0;
      acc += x;
    }
    __syncthreads();
  }
  if (acc == -1.0) sink[0] = acc;
}

}  // namespace yalps
