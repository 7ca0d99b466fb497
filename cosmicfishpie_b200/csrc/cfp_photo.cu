// Photometric 3x2pt: Limber kernels, C_ell contraction on FP64 tensor cores, Cholesky/trace Fisher.
//
// Kernels (all FP64):
//   photo_sqrtp_kernel    T[c][l][z] = sqrt(P(z,(l+1/2)/chi))/chi for kappa < kmax, else 0; point-wise
//                         bicubic B-spline evaluation with FITPACK semantics        (photo_obs.py:457-512)
//   photo_eff_kernel      lensing efficiency by backward cumulative trapezoid      (photo_obs.py:113-160)
//   photo_window_kernel   per stencil point: W_a(z)/sqrt(H) and W^IA_a(z)/sqrt(H)   (photo_obs.py:514-722)
//   photo_cls_kernel      C_l[a][b] = sum_z w_z K_a K_b, K_a = T W_a + T W^IA_a, as a batch of
//                         (Nb x Nz)(Nz x Nb) SYRKs on mma.sync m8n8k4 f64 (DMMA)    (photo_obs.py:878-928)
//   photo_fisher_kernel   per multipole bin: M_p = Phi dC_p streamed from the stencil's C_ell, Gauss-Jordan inverse
//                         of the covariance, Z_p = A^-1 M_p and F_pq += w_l Tr(Z_p Z_q) on DMMA
//                         (photo_cov.py:198-245, cosmicfish.py:643-744)
//   photo_reduce_kernel   fixed-tree sum over multipole bins
//   photo_chi2_reg_kernel likelihood: one warp per (point, multipole), register-resident Cholesky fused with the forward substitution,
//                         tr(A^-1 B) = |L^-1 L_B|^2                                      (photo_like.py:28-97,180-256)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

#include "cfp_common.cuh"

namespace {

constexpr int CLS_WARPS = 8;  // warps per CTA (A/B on B200, CFP_CLS_WARPS: Fisher job 4: 0.086, 8: 0.076, 16: 0.082 ms;
                              // 4096-point likelihood chunk incl. chi2 4: 7.83, 8: 7.37, 16: 8.34 ms)
constexpr int MAX_PNB = 8;  // Nb <= 64 tomographic rows

// multipoles [l0, l0 + nl) only: a rank of a sharded forecast fills its own rows of the table
__global__ void photo_sqrtp_kernel(const double* __restrict__ blob, const int64_t* __restrict__ desc, int n_tables,
                                   int tab, int NC, int Nl, int Nz, int l0, int nl, const double* __restrict__ ell,
                                   const double* __restrict__ z, const double* __restrict__ chi, double kmax,
                                   double* __restrict__ T) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tot = (size_t)NC * nl * Nz;
  if (idx >= tot) return;
  const int iz = (int)(idx % Nz);
  const int il = l0 + (int)((idx / Nz) % nl);
  const int c = (int)(idx / ((size_t)Nz * nl));
  const double ch = chi[(size_t)c * Nz + iz];
  const double kappa = (ell[il] + 0.5) / ch;  // photo_obs.py:481
  double v = 0.0;
  if (kappa < kmax) {
    const int64_t* d = desc + ((size_t)c * n_tables + tab) * CFP_BANK_DESC;
    const double p = cfp_bispev_point(blob + d[2], (int)d[0], blob + d[3], (int)d[1], blob + d[4], z[iz], kappa);
    v = sqrt(p) / ch;
  }
  T[((size_t)c * Nl + il) * Nz + iz] = v;
}

// one CTA per cosmology: eff = I1 - chi * I2 with backward cumulative trapezoids (sequential in z like
// np.cumsum).  n_i(z) and n_i(z)/chi(z) are staged in shared memory by the whole CTA first, so the serial
// 200-step recurrence of each lensing row runs on shared-memory operands only.
__global__ void __launch_bounds__(256) photo_eff_kernel(int NP, int Nb, int Nz, const int* __restrict__ kind,
                                                        const int* __restrict__ pair_c, const int* __restrict__ pair_w,
                                                        const double* __restrict__ ngal_all,
                                                        const double* __restrict__ chi, double dz,
                                                        double* __restrict__ eff) {
  extern __shared__ double effsm[];  // chi[Nz], n[Nb][Nz], q[Nb][Nz], e[Nb][Nz]
  // one CTA per (cosmology, n(z) window set) pair that occurs among the stencil points
  const int c = pair_c ? pair_c[blockIdx.x] : blockIdx.x;
  const double* ngal = ngal_all + (size_t)(pair_w ? pair_w[blockIdx.x] : 0) * Nb * Nz;
  double* sch = effsm;
  double* sn = sch + Nz;
  double* sq = sn + (size_t)Nb * Nz;
  double* se = sq + (size_t)Nb * Nz;
  for (int i = threadIdx.x; i < Nz; i += blockDim.x) sch[i] = chi[(size_t)c * Nz + i];
  __syncthreads();
  for (int i = threadIdx.x; i < Nb * Nz; i += blockDim.x) {
    const double v = ngal[i];
    sn[i] = v;
    sq[i] = v / sch[i % Nz];
  }
  __syncthreads();
  // trapezoid increments of both integrals, computed by the whole CTA: sq <- t2, se <- t1
  for (int i = threadIdx.x; i < Nb * Nz; i += blockDim.x)
    se[i] = (i % Nz < Nz - 1) ? 0.5 * (sq[i + 1] + sq[i]) * dz : 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < Nb * Nz; i += blockDim.x) sq[i] = se[i];
  __syncthreads();
  for (int i = threadIdx.x; i < Nb * Nz; i += blockDim.x)
    se[i] = (i % Nz < Nz - 1) ? 0.5 * (sn[i + 1] + sn[i]) * dz : 0.0;
  __syncthreads();
  // the serial part (np.cumsum order, photo_obs.py:137-147) is two additions and one FMA per step
  for (int a = threadIdx.x; a < Nb; a += blockDim.x) {
    double* e = se + (size_t)a * Nz;
    if (kind[a] != 0) {
      for (int i = 0; i < Nz; ++i) e[i] = 0.0;
      continue;
    }
    const double* t2 = sq + (size_t)a * Nz;
    double I1 = 0.0, I2 = 0.0;
    e[Nz - 1] = I1 - sch[Nz - 1] * I2;
#pragma unroll 4
    for (int i = Nz - 2; i >= 0; --i) {
      I1 = I1 + e[i];
      I2 = I2 + t2[i];
      e[i] = I1 - sch[i] * I2;
    }
  }
  __syncthreads();
  double* out = eff + (size_t)blockIdx.x * Nb * Nz;
  for (int i = threadIdx.x; i < Nb * Nz; i += blockDim.x) out[i] = se[i];
}

// fallback for grids whose rows do not fit in shared memory: one thread per (cosmology, row), global operands
__global__ void photo_eff_kernel_global(int NP, int Nb, int Nz, const int* __restrict__ kind,
                                        const int* __restrict__ pair_c, const int* __restrict__ pair_w,
                                        const double* __restrict__ ngal_all, const double* __restrict__ chi, double dz,
                                        double* __restrict__ eff) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NP * Nb) return;
  const int pr = idx / Nb, a = idx % Nb;
  const int c = pair_c ? pair_c[pr] : pr;
  const double* ngal = ngal_all + (size_t)(pair_w ? pair_w[pr] : 0) * Nb * Nz;
  double* e = eff + (size_t)idx * Nz;
  if (kind[a] != 0) {
    for (int i = 0; i < Nz; ++i) e[i] = 0.0;
    return;
  }
  const double* n = ngal + (size_t)a * Nz;
  const double* ch = chi + (size_t)c * Nz;
  double I1 = 0.0, I2 = 0.0;
  e[Nz - 1] = I1 - ch[Nz - 1] * I2;
  double n_hi = n[Nz - 1], q_hi = n[Nz - 1] / ch[Nz - 1];
  for (int i = Nz - 2; i >= 0; --i) {
    const double n_lo = n[i], q_lo = n[i] / ch[i];
    I1 = I1 + 0.5 * (n_hi + n_lo) * dz;
    I2 = I2 + 0.5 * (q_hi + q_lo) * dz;
    e[i] = I1 - ch[i] * I2;
    n_hi = n_lo;
    q_hi = q_lo;
  }
}

// U[s][a][z] = W_a/sqrt(H), V[s][a][z] = W^IA_a/sqrt(H), rows padded to ldz
__global__ void photo_window_kernel(int S, int Nb, int Nz, int ldz, int nrow_pad, const int* __restrict__ cosmo_of_s,
                                    const int* __restrict__ win_of_s, const int* __restrict__ pair_of_s,
                                    const int* __restrict__ kind, const double* __restrict__ z,
                                    const double* __restrict__ chi, const double* __restrict__ hub,
                                    const double* __restrict__ wlpref, const double* __restrict__ ngal,
                                    const double* __restrict__ bias, int Kb, int bias_rows,
                                    const int* __restrict__ zmap, const double* __restrict__ ia,
                                    const double* __restrict__ eff, double* __restrict__ U, double* __restrict__ V) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tot = (size_t)S * nrow_pad * ldz;
  if (idx >= tot) return;
  const int iz = (int)(idx % ldz);
  const int a = (int)((idx / ldz) % nrow_pad);
  const int s = (int)(idx / ((size_t)ldz * nrow_pad));
  double u = 0.0, v = 0.0;
  if (a < Nb && iz < Nz) {
    const int c = cosmo_of_s[s];
    const double H = hub[(size_t)c * Nz + iz];
    const double sh = sqrt(H);
    const double ng = ngal[((size_t)(win_of_s ? win_of_s[s] : 0) * Nb + a) * Nz + iz];
    const int pr = pair_of_s ? pair_of_s[s] : c;  // (cosmology, window set) pair: row of the efficiency table
    if (kind[a] == 0) {  // lensing row, photo_obs.py:584-590
      const double w = wlpref[c] * (1.0 + z[iz]) * chi[(size_t)c * Nz + iz] * eff[((size_t)pr * Nb + a) * Nz + iz];
      const double wia = ng * ia[(size_t)s * Nz + iz] * H;
      u = w / sh;
      v = wia / sh;
    } else {  // clustering row, photo_obs.py:542, 715-717
      const double w = (ng * H) * bias[((size_t)s * bias_rows + (bias_rows == 1 ? 0 : a)) * Kb + (zmap ? zmap[iz] : iz)];
      u = w / sh;
      v = 0.0;
    }
  }
  U[idx] = u;
  V[idx] = v;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct ClsParams {
  int S, Nl, Nz, Nb, nb, ldz, nz4, lch, l_begin, l_end;
  const int* cosmo_of_s;
  const double *U, *V;   // [S][nb*8][ldz]
  const double* T;       // [NC][Nl][Nz] sqrt(P)/chi of the lensing rows
  const double* Tg;      // the same for the clustering rows (another table when GCph_Tracer = 'clustering', else = T)
  const int* kind;       // [Nb] 0 = lensing row, 1 = clustering row
  const double* wz;      // [nz4] trapezoid weights (zero padded)
  const double *ell, *lmin, *lmax;
  double* cls;           // [S][Nl][Nb][Nb]
};

template <int NB, bool TWO, int WARPS>  // TWO: the clustering rows use their own sqrt(P) table (GCph_Tracer = 'clustering')
__global__ void __launch_bounds__(WARPS * 32) photo_cls_kernel(ClsParams P) {
  constexpr int NT = NB * (NB + 1) / 2;
  extern __shared__ double smem[];
  const int rows = NB * 8;
  double* sK = smem;  // [rows][ldz]: W + W^IA of the point (both multiply the same sqrt(P)/chi, photo_obs.py:905-919);
                      // one staged array instead of two doubles the resident CTAs per SM
  const int s = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const double* gU = P.U + (size_t)s * rows * P.ldz;
    const double* gV = P.V + (size_t)s * rows * P.ldz;
    for (int i = tid; i < rows * P.ldz; i += blockDim.x) sK[i] = gU[i] + gV[i];
  }
  __syncthreads();
  const int c = P.cosmo_of_s[s];
  const int l0 = P.l_begin + blockIdx.x * P.lch;
  const int l1 = min(l0 + P.lch, P.l_end);
  const int g = lane >> 2, tg = lane & 3;
  // which sqrt(P) table each of this lane's rows multiplies (photo_obs.py:483-507: GCph uses its tracer's spectrum)
  bool gc_row[NB];
#pragma unroll
  for (int ib = 0; ib < NB; ++ib) {
    const int row = ib * 8 + g;
    gc_row[ib] = TWO && row < P.Nb && P.kind[row] != 0;
  }
  for (int il = l0 + warp; il < l1; il += (int)(blockDim.x >> 5)) {
    const double* Trow = P.T + ((size_t)c * P.Nl + il) * P.Nz;
    const double* Tgrow = P.Tg + ((size_t)c * P.Nl + il) * P.Nz;
    double acc[NT][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = 0.0;
    for (int z0 = 0; z0 < P.nz4; z0 += 4) {
      const int iz = z0 + tg;
      const double t = (iz < P.Nz) ? __ldg(Trow + iz) : 0.0;
      const double t_gc = TWO ? ((iz < P.Nz) ? __ldg(Tgrow + iz) : 0.0) : t;
      const double w = P.wz[iz];
      double a[NB];
#pragma unroll
      for (int ib = 0; ib < NB; ++ib) {
        const int off = (ib * 8 + g) * P.ldz + iz;
        a[ib] = ((TWO && gc_row[ib]) ? t_gc : t) * sK[off];
      }
      int tt = 0;
#pragma unroll
      for (int ib = 0; ib < NB; ++ib)
#pragma unroll
        for (int jb = 0; jb <= ib; ++jb) {
          dmma884(acc[tt][0], acc[tt][1], a[ib], w * a[jb]);
          ++tt;
        }
    }
    // masked write-back (both triangles); rows outside their multipole range carry no signal
    const double L = P.ell[il];
    double* out = P.cls + ((size_t)s * P.Nl + il) * P.Nb * P.Nb;
    int tt = 0;
#pragma unroll
    for (int ib = 0; ib < NB; ++ib)
#pragma unroll
      for (int jb = 0; jb <= ib; ++jb) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int i = ib * 8 + g, j = jb * 8 + tg * 2 + r;
          if (i < P.Nb && j < P.Nb && j <= i) {
            const bool ok = (L >= P.lmin[i]) && (L <= P.lmax[i]) && (L >= P.lmin[j]) && (L <= P.lmax[j]);
            const double v = ok ? acc[tt][r] : 0.0;
            out[(size_t)i * P.Nb + j] = v;
            out[(size_t)j * P.Nb + i] = v;
          }
        }
        ++tt;
      }
  }
}

struct FishParams {
  int S, Nl, Nb, npar, l_begin, l_end, ldb, zstride;
  const double* cls;  // [S][Nl][Nb][Nb]
  const double *ell, *noise, *sqrtfsky;
  const int* stp_ptr;   // CSR of the stencil per parameter: stencil points and weights
  const int* stp_s;
  const double* stp_w;
  const double* stden;  // [npar]
  const double* strden; // [npar] 1 / stden
  double* Y;          // scratch [Nl-1][npar][zstride]
  double* Fl;         // [Nl-1][npar][npar]
};

// One CTA (16 warps) per multipole bin (left-edge index li).  All steps are data-parallel:
//   1. M_p = Phi dC_p for every parameter, streamed from the stencil's C_ell (one warp per (parameter, 8 rows),
//      16 independent coalesced loads in flight per lane) straight into the buffer Z_p will replace
//   2. A = C_l(fiducial) + N in shared memory, inverted in place by Gauss-Jordan (SPD, no pivoting needed;
//      cond <= 2e4, the reference itself uses an explicit pinv)
//   3. Z_p = A^-1 M_p in place on the FP64 tensor cores (mma.m8n8k4), one warp per (parameter, 8 columns)
//   4. F_pq = w_l sum_ij Z_p[i][j] Z_q[j][i] = w_l Tr(dC_p Sigma^-1 dC_q Sigma^-1) as a (npar x K)(K x npar) product
//      over K = (i, j) on the tensor cores; warps split K, partial tiles are added in fixed warp order
// The Z_p live in shared memory when they fit (ZSMEM), else in a global scratch.
constexpr int FISH_THREADS = 512;
constexpr int FISH_WARPS = FISH_THREADS / 32;
constexpr int FISH_MAX_PT = 8;   // npar <= 64

template <int KTMAX, bool ZSMEM>  // KTMAX: k-steps of step 3 the kernel is unrolled for (Nb <= 4 KTMAX)
__global__ void __launch_bounds__(FISH_THREADS) photo_fisher_kernel(FishParams P) {
  extern __shared__ double sm[];
  const int Nb = P.Nb, ld = P.ldb;
  const int PT = (P.npar + 7) / 8;
  double* sA = sm;                                   // [Nb][ld]
  double* sR = sA + Nb * ld;                         // [FISH_WARPS][PT][64] cross-warp reduction of step 4
  double* sZ = sR + FISH_WARPS * PT * 64;            // [npar][zstride] if ZSMEM
  const int li = P.l_begin + blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (li >= P.l_end) return;
  const int ps = P.zstride;  // doubles between consecutive Z_p (padded: conflict-free fragment loads below)
  double* Z = ZSMEM ? sZ : (P.Y + (size_t)blockIdx.x * P.npar * ps);
  const double* C0 = P.cls + ((size_t)0 * P.Nl + li) * Nb * Nb;
  for (int e = tid; e < Nb * Nb; e += FISH_THREADS) {
    const int i = e / Nb, j = e - i * Nb;
    sA[i * ld + j] = C0[e] + ((i == j) ? P.noise[i] : 0.0);
  }
  // ---- 1. M_p = Phi dC_p (photo_cov.py:228); padding columns are zero ------------------------------------
  const int RB = (Nb + 7) / 8;
  for (int unit = warp; unit < P.npar * RB; unit += FISH_WARPS) {
    const int p = unit / RB, i0 = (unit - p * RB) * 8;
    const int q0 = P.stp_ptr[p], nq = P.stp_ptr[p + 1] - q0;
    const double rden = P.strden[p];
    double* Zp = Z + (size_t)p * ps;
    if (nq <= 4) {
      const double* src[4];
      double wq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool on = u < nq;
        src[u] = P.cls + ((size_t)(on ? P.stp_s[q0 + u] : 0) * P.Nl + li) * Nb * Nb;
        wq[u] = on ? P.stp_w[q0 + u] : 0.0;
      }
      for (int j = lane; j < ld; j += 32) {
        const bool jok = j < Nb;
        const int jj = jok ? j : Nb - 1;
#pragma unroll
        for (int ih = 0; ih < 8; ih += 4) {
          double v[4][4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int ii = min(i0 + ih + r, Nb - 1);
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u][r] = (u < nq) ? __ldg(src[u] + (size_t)ii * Nb + jj) : 0.0;
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int i = i0 + ih + r;
            double acc = 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += wq[u] * v[u][r];
            if (i < Nb) Zp[i * ld + j] = jok ? (acc * rden) * P.sqrtfsky[i] : 0.0;
          }
        }
      }
    } else {
      for (int j = lane; j < ld; j += 32)
        for (int i = i0; i < min(i0 + 8, Nb); ++i) {
          double acc = 0.0;
          if (j < Nb)
            for (int q = q0; q < q0 + nq; ++q)
              acc += P.stp_w[q] * __ldg(P.cls + (((size_t)P.stp_s[q] * P.Nl + li) * Nb + i) * Nb + j);
          Zp[i * ld + j] = (acc * rden) * P.sqrtfsky[i];
        }
    }
  }
  __syncthreads();
  // ---- 2. in-place Gauss-Jordan inversion: every thread computes the new values of its (<= 8) elements from the
  // old matrix into registers, the CTA synchronises, then they are written back
  {
    const int per = (Nb * Nb + FISH_THREADS - 1) / FISH_THREADS;  // uniform across the CTA (Nb <= 64 -> per <= 8)
    int eo[8], ec[8], er[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int e = tid + it * FISH_THREADS;
      const int i = e / Nb, j = e - i * Nb;
      er[it] = (e < Nb * Nb) ? i : -1;
      ec[it] = j;
      eo[it] = i * ld + j;
    }
    for (int k = 0; k < Nb; ++k) {
      const double piv = 1.0 / sA[k * ld + k];
      double v[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        if (it < per && er[it] >= 0) {
          const int i = er[it], j = ec[it];
          const double aik = sA[i * ld + k], akj = sA[k * ld + j];
          const double t = akj * piv;
          double r = sA[eo[it]] - aik * t;
          r = (j == k) ? -aik * piv : r;
          r = (i == k) ? ((j == k) ? piv : t) : r;
          v[it] = r;
        }
      }
      __syncthreads();
#pragma unroll
      for (int it = 0; it < 8; ++it)
        if (it < per && er[it] >= 0) sA[eo[it]] = v[it];
      __syncthreads();
    }
  }
  // ---- 3. Z_p = A^-1 M_p in place: one warp per (parameter, block of 8 columns) -----------------------
  const int g = lane >> 2, tg = lane & 3;  // mma.m8n8k4 fragment coordinates
  const int MT = (Nb + 7) / 8, KT = (Nb + 3) / 4;
  for (int unit = warp; unit < P.npar * MT; unit += FISH_WARPS) {
    const int p = unit / MT, nt = unit - p * MT;
    double* Zp = Z + (size_t)p * ps;
    const int jb = nt * 8 + g;
    double bf[KTMAX];  // the 8 columns of M_p this warp replaces: all k-steps of the B operand
#pragma unroll
    for (int kt = 0; kt < KTMAX; ++kt) {
      const int k = kt * 4 + tg;
      bf[kt] = (kt < KT && k < Nb && jb < Nb) ? Zp[k * ld + jb] : 0.0;
    }
    __syncwarp();
    for (int mt = 0; mt < MT; ++mt) {
      const int i = mt * 8 + g;
      const bool iok = i < Nb;
      const double* arow = sA + (iok ? i : 0) * ld + tg;
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int kt = 0; kt < KTMAX; ++kt) {
        const double av = (kt < KT && iok && kt * 4 + tg < Nb) ? arow[kt * 4] : 0.0;
        if (kt < KT) dmma884(c0, c1, av, bf[kt]);
      }
      const int j = nt * 8 + tg * 2;
      if (iok && j < ld) Zp[i * ld + j] = c0;
      if (iok && j + 1 < ld) Zp[i * ld + j + 1] = c1;
    }
  }
  __syncthreads();
  // ---- 4. F_l[p][q] = w_l sum_ij Z_p[i][j] Z_q[j][i]  (cosmicfish.py:681-683, 704, 723) -------------------
  const double lc = 0.5 * P.ell[li] + 0.5 * P.ell[li + 1];
  const double wl = (lc + 0.5) * (P.ell[li + 1] - P.ell[li]);
  double* Fl = P.Fl + (size_t)blockIdx.x * P.npar * P.npar;
  const int KS = (Nb * ld) / 4;  // ld is a multiple of 4: a k-step never straddles two rows i
  // B operand: B[k = (i, j)][q] = Z_q[j][i]; per-lane bases of the (at most FISH_MAX_PT) parameter blocks
  const double* zq[FISH_MAX_PT];
#pragma unroll
  for (int u = 0; u < FISH_MAX_PT; ++u) {
    const int qb = u * 8 + g;
    zq[u] = Z + (size_t)(qb < P.npar ? qb : 0) * ps;
  }
  for (int pt = 0; pt < PT; ++pt) {
    const int pa = pt * 8 + g;
    const bool pok = pa < P.npar;
    const double* Za = Z + (size_t)(pok ? pa : 0) * ps + tg;
    double c[FISH_MAX_PT][2];
#pragma unroll
    for (int u = 0; u < FISH_MAX_PT; ++u) c[u][0] = c[u][1] = 0.0;
    // k-step ks covers elements e = 4 ks + tg of row i = e / ld; this warp takes ks = warp, warp + 16, ...
    int i = (warp * 4) / ld, j0 = warp * 4 - i * ld;
    for (int ks = warp; ks < KS; ks += FISH_WARPS) {
      const int j = j0 + tg;
      const double av = pok ? Za[ks * 4] : 0.0;
      const bool jok = j < Nb;
      const int off = (jok ? j : 0) * ld + i;
#pragma unroll
      for (int u = 0; u < FISH_MAX_PT; ++u) {
        if (u <= pt) {
          const double bv = (jok && u * 8 + g < P.npar) ? zq[u][off] : 0.0;
          dmma884(c[u][0], c[u][1], av, bv);
        }
      }
      j0 += FISH_WARPS * 4;
      while (j0 >= ld) {
        j0 -= ld;
        ++i;
      }
    }
    // fixed-order reduction over the warps
    __syncthreads();
#pragma unroll
    for (int u = 0; u < FISH_MAX_PT; ++u)
      if (u <= pt) {
        sR[((size_t)warp * PT + u) * 64 + lane * 2 + 0] = c[u][0];
        sR[((size_t)warp * PT + u) * 64 + lane * 2 + 1] = c[u][1];
      }
    __syncthreads();
    for (int x = tid; x < (pt + 1) * 64; x += FISH_THREADS) {
      const int u = x / 64, l = (x % 64) / 2, r = x % 2;
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < FISH_WARPS; ++w) sum += sR[((size_t)w * PT + u) * 64 + l * 2 + r];
      const int pi = pt * 8 + (l >> 2), qi = u * 8 + (l & 3) * 2 + r;
      if (pi < P.npar && qi < P.npar && qi <= pi) {
        sum *= wl;
        Fl[pi * P.npar + qi] = sum;
        Fl[qi * P.npar + pi] = sum;
      }
    }
  }
}

// F = sum over multipole bins in a FIXED tree: one warp per element, lane l adds bins l, l+32, ... in order, then a
// butterfly over the lanes.  All bins of an element are fetched concurrently instead of one dependent chain.
__global__ void __launch_bounds__(256) photo_reduce_kernel(const double* __restrict__ Fl, int nbins, int n,
                                                           double* __restrict__ F) {
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= n) return;
  double s = 0.0;
  for (int l = lane; l < nbins; l += 32) s += Fl[(size_t)l * n + e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) F[e] = s;
}


// ---------------------------------------------------------------------------- likelihood
__device__ __forceinline__ double chi_rsqrt(double x) {  // x > 0, normal: 1/sqrt(x) to ~1 ulp without the slow path
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  return fma(y, e, y);
}

// One warp per (point, multipole) takes every block of the chi^2 in turn.  Lane j owns ROW j of the (lower) matrix
// A = C_l + N and COLUMN j of the data factor L_B, both in registers (static indices only: the loops are unrolled
// for a compile-time bound N >= n).  Right-looking Cholesky: at step k the pivot is broadcast by a shuffle, every
// lane scales its element of column k, the column goes through a 2 x 32-double shared-memory buffer (one store,
// then 128-bit broadcast loads) and is used TWICE per loaded value: for the trailing update of A and for the forward
// substitution M = L^-1 L_B, which is fused into the same sweep -- tr(A^-1 B) = |M|_F^2 accumulates as the columns
// are finalised.  No pair lists, one warp barrier per pivot.
// A block that is the LEADING sub-block of another one (3x2pt: the lensing block inside the full matrix) costs
// nothing: the Cholesky factor, L_B and M of a leading block are the leading blocks of the big ones, so its
// tr and ln det are partial sums of the same sweep.
struct ChiBlocks {
  int nblk;
  int off[8], n[8];
  int parent[8];  // block whose leading sub-block this one is (-1: none)
  int child[8];   // the block that is this one's leading sub-block (-1: none)
};

template <int N, bool DATA, bool EXACT = false>  // EXACT: n == N, every size test folds at compile time
__device__ __forceinline__ void chol_trace(const double* __restrict__ A, int lda, const double* __restrict__ noise,
                                           const double* __restrict__ LB, int ldb, int n, int nlead,
                                           double* col /* not __restrict__: the barriers must order its loads */,
                                           int lane, double& tr, double& tr_lead, double& lndet, double& lndet_lead,
                                           double* __restrict__ LBout) {
  double a[N];
  if (EXACT) n = N;
  const bool in = lane < n;
  const double nz = (noise && in) ? noise[lane] : 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i)  // row `lane` of the symmetric matrix read as column `lane`: coalesced; beyond n: identity
    a[i] = ((EXACT || i < n) && in) ? A[(size_t)i * lda + lane] + (i == lane ? nz : 0.0) : (i == lane ? 1.0 : 0.0);
  double diag = 1.0;
  // ---- pass 1: right-looking Cholesky; column k (scaled) is left in shared memory col[k][0..31], with 1 / L_kk in
  // its diagonal slot
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (EXACT || k < n) {  // (uniform)
      const double akk = __shfl_sync(0xffffffffu, a[k], k);
      const double rd = chi_rsqrt(akk);
      const double ljk = a[k] * rd;  // lane j > k: L[j][k]; lane k: L[k][k]
      a[k] = ljk;
      if (lane == k) diag = ljk;
      double* cb = col + k * 32;
      cb[lane] = (lane == k) ? rd : ljk;
      __syncwarp();
      const double nl = -ljk;
      int i = k + 1;
      if (i & 1) {
        if (i < N) a[i] = fma(nl, cb[i], a[i]);
        ++i;
      }
#pragma unroll
      for (; i + 1 < N; i += 2) {
        const double2 c = *reinterpret_cast<const double2*>(cb + i);
        a[i] = fma(nl, c.x, a[i]);
        a[i + 1] = fma(nl, c.y, a[i + 1]);
      }
      if ((N & 1) && i < N) a[i] = fma(nl, cb[i], a[i]);  // (odd N: the last element has no partner)
    }
  }
  const double lg = in ? log(diag) : 0.0;
  double l_all = lg, l_lead = (lane < nlead) ? lg : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l_all += __shfl_xor_sync(0xffffffffu, l_all, o);
    l_lead += __shfl_xor_sync(0xffffffffu, l_lead, o);
  }
  lndet = 2.0 * l_all;
  lndet_lead = 2.0 * l_lead;
  tr = tr_lead = 0.0;
  if constexpr (DATA) {
    if (in) {  // keep the factor: row `lane`, upper triangle zero
#pragma unroll
      for (int k = 0; k < N; ++k)
        if (EXACT || k < n) LBout[(size_t)lane * ldb + k] = (k <= lane) ? a[k] : 0.0;
    }
  } else {
  // ---- pass 2: M = L^-1 L_B by columns (lane c owns column c of L_B, zero above the diagonal), the columns of L
  // read back from shared memory; tr(A^-1 B) = |M|_F^2 accumulates as the rows are finalised
#pragma unroll
  for (int i = 0; i < N; ++i) a[i] = ((EXACT || i < n) && in) ? LB[(size_t)i * ldb + lane] : 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (!EXACT && k >= n) break;
    {
      __syncwarp();  // (also keeps the scheduler from hoisting the broadcast loads of every later step up here)
      const double* cb = col + k * 32;
      const double xk = a[k] * cb[k];  // M[k][lane], final
      const double x2 = xk * xk;
      tr += x2;
      if (k < nlead && lane < nlead) tr_lead += x2;
      const double nx = -xk;
      int i = k + 1;
      if (i & 1) {
        if (i < N) a[i] = fma(nx, cb[i], a[i]);
        ++i;
      }
#pragma unroll
      for (; i + 1 < N; i += 2) {
        const double2 c = *reinterpret_cast<const double2*>(cb + i);
        a[i] = fma(nx, c.x, a[i]);
        a[i + 1] = fma(nx, c.y, a[i + 1]);
      }
      if ((N & 1) && i < N) a[i] = fma(nx, cb[i], a[i]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tr += __shfl_xor_sync(0xffffffffu, tr, o);
    tr_lead += __shfl_xor_sync(0xffffffffu, tr_lead, o);
  }
  }
  __syncwarp();  // (the next block of this warp overwrites col)
}

// NQ: the kernel is compiled for blocks of at most 4 NQ rows (registers: the largest instantiation counts)
template <bool DATA, int NQ>
__device__ __forceinline__ void chol_dispatch(int n, const double* A, int lda, const double* noise, const double* LB,
                                              int ldb, int nlead, double* col, int lane, double& tr, double& tr_lead,
                                              double& lndet, double& lndet_lead, double* LBout) {
#define CHOL_CASE(Q)                                                                                               \
  if constexpr (Q < NQ) {                                                                                          \
    if (n <= 4 * Q) {                                                                                              \
      chol_trace<4 * Q, DATA>(A, lda, noise, LB, ldb, n, nlead, col, lane, tr, tr_lead, lndet, lndet_lead, LBout); \
      return;                                                                                                      \
    }                                                                                                              \
  }
  // the Euclid tomography (13 + 13 bins) in exact sizes: straight-line code, no padded pivots
  if constexpr (4 * NQ >= 26) {
    if (n == 26) {
      chol_trace<26, DATA, true>(A, lda, noise, LB, ldb, n, nlead, col, lane, tr, tr_lead, lndet, lndet_lead, LBout);
      return;
    }
  }
  if constexpr (4 * NQ >= 13) {
    if (n == 13) {
      chol_trace<13, DATA, true>(A, lda, noise, LB, ldb, n, nlead, col, lane, tr, tr_lead, lndet, lndet_lead, LBout);
      return;
    }
  }
  CHOL_CASE(1)
  CHOL_CASE(2)
  CHOL_CASE(3)
  CHOL_CASE(4)
  CHOL_CASE(5)
  CHOL_CASE(6)
  CHOL_CASE(7)
#undef CHOL_CASE
  chol_trace<4 * NQ, DATA>(A, lda, noise, LB, ldb, n, nlead, col, lane, tr, tr_lead, lndet, lndet_lead, LBout);
}

constexpr int CHR_WARPS = 4;
// q[s][l][b] = tr(A^-1 B) + ln det A - ln det B - n  (photo_like.py:28-97); DATA: q[l][b] = ln det B, LB = chol(B)
// resident CTAs per SM the kernel is compiled for: caps the registers near 2 N + 40 (without a cap the scheduler hoists
// the broadcast loads of the whole substitution pass and takes 255)
constexpr int chr_min_blocks(int nq) { return nq <= 2 ? 8 : (nq <= 4 ? 7 : (nq <= 5 ? 5 : 4)); }
template <bool DATA, int NQ>
__global__ void __launch_bounds__(CHR_WARPS * 32, chr_min_blocks(NQ)) photo_chi2_reg_kernel(
    int S, int Nl, int Nb, ChiBlocks B, const double* __restrict__ cls, const double* __restrict__ noise,
    const double* __restrict__ wts, double* __restrict__ LB, const double* __restrict__ lndetB, int nmax,
    double* __restrict__ q) {
  __shared__ __align__(16) double colbuf[CHR_WARPS][4 * NQ * 32];  // per warp: the columns of L, [k][row]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t unit = (size_t)blockIdx.x * CHR_WARPS + warp;
  if (unit >= (size_t)S * Nl) return;
  const int il = (int)(unit % Nl), s = (int)(unit / Nl);
  const double* Afull = cls + ((size_t)s * Nl + il) * Nb * Nb;
  for (int b = 0; b < B.nblk; ++b) {
    const size_t qi = ((size_t)s * Nl + il) * B.nblk + b;
    const bool live = wts[(size_t)b * Nl + il] != 0.0;  // inside the block's multipole range
    if (!live) {
      if (lane == 0) q[qi] = 0.0;
      continue;
    }
    const int off = B.off[b], n = B.n[b];
    double* LBu = LB + ((size_t)il * B.nblk + b) * nmax * nmax;
    double tr, trl, ld, ldl;
    if (DATA) {
      chol_dispatch<true, NQ>(n, Afull + (size_t)off * Nb + off, Nb, nullptr, nullptr, nmax, 0, colbuf[warp], lane, tr, trl,
                              ld, ldl, LBu);
      if (lane == 0) q[qi] = ld;
      continue;
    }
    const int par = B.parent[b];
    if (par >= 0 && wts[(size_t)par * Nl + il] != 0.0) continue;  // written together with the parent
    int ch = B.child[b];
    if (ch >= 0 && wts[(size_t)ch * Nl + il] == 0.0) ch = -1;
    chol_dispatch<false, NQ>(n, Afull + (size_t)off * Nb + off, Nb, noise + off, LBu, nmax, ch >= 0 ? B.n[ch] : 0,
                             colbuf[warp], lane, tr, trl, ld, ldl, nullptr);
    if (lane == 0) {
      q[qi] = tr + ld - lndetB[(size_t)il * B.nblk + b] - (double)n;
      if (ch >= 0) q[qi - b + ch] = trl + ldl - lndetB[(size_t)il * B.nblk + ch] - (double)B.n[ch];
    }
  }
}

// parent / child maps of the blocks: d is a leading sub-block of c if they start at the same row and d is smaller
static ChiBlocks chi_blocks(int nblk, const int32_t* off, const int32_t* n) {
  ChiBlocks B;
  B.nblk = nblk;
  for (int b = 0; b < 8; ++b) B.off[b] = B.n[b] = 0, B.parent[b] = B.child[b] = -1;
  for (int b = 0; b < nblk; ++b) B.off[b] = off[b], B.n[b] = n[b];
  for (int d = 0; d < nblk; ++d)
    for (int c = 0; c < nblk; ++c)
      if (c != d && off[c] == off[d] && n[c] > n[d] && B.child[c] < 0 && B.parent[d] < 0 && B.parent[c] < 0) {
        B.parent[d] = c;
        B.child[c] = d;
      }
  return B;
}

// data pass (ln det B, L_B) + theory pass + multipole sum
static cudaError_t photo_launch_chi2(cfp_ctx* ctx, cudaStream_t st, int S, int Nl, int Nb, int nblk, const int32_t* off_h,
                                     const int32_t* n_h, int nmax, const double* theory_d, const double* noise_d,
                                     const double* data_d, const double* w_d, double* lb_d, double* lndet_d, double* q_d,
                                     double* chi2_d);

// ---- Fisher matrix from caller-supplied arrays (the reference's seam photo_LSS_fishermatrix_einsum(noisy_cls,
// covmat, derivs), cosmicfish.py:643-744): per multipole bin l (left edge), A = covmat[l], Z_p = A^-1 dC_p[l],
// F_pq(l) = w_l Tr(Z_p Z_q), w_l = (l_centre + 1/2) (l_{i+1} - l_i).  One CTA per bin; A^-1 by Gauss-Jordan with
// partial pivoting in shared memory (the reference takes the pseudo-inverse; the matrices are non-singular), the Z_p
// of all parameters in global scratch.  Not a hot path: the forecast itself runs photo_fisher_kernel on the stencil.
__global__ void __launch_bounds__(256) photo_fisher_user_kernel(int Nl, int Nb, int npar, const double* __restrict__ cov,
                                                                const double* __restrict__ dC, const double* __restrict__ ell,
                                                                double* __restrict__ Zs, double* __restrict__ Fl) {
  extern __shared__ double fsm[];
  const int l = blockIdx.x, tid = threadIdx.x, n = Nb, ld = 2 * n;
  double* aug = fsm;  // [n][2n]: (A | I) -> (I | A^-1)
  __shared__ int piv;
  for (int e = tid; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    aug[i * ld + j] = cov[((size_t)l * n + i) * n + j];
    aug[i * ld + n + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    if (tid == 0) {
      int best = k;
      double bv = fabs(aug[k * ld + k]);
      for (int i = k + 1; i < n; ++i)
        if (fabs(aug[i * ld + k]) > bv) bv = fabs(aug[i * ld + k]), best = i;
      piv = best;
    }
    __syncthreads();
    if (piv != k)
      for (int j = tid; j < ld; j += blockDim.x) {
        const double t = aug[k * ld + j];
        aug[k * ld + j] = aug[piv * ld + j];
        aug[piv * ld + j] = t;
      }
    __syncthreads();
    const double d = 1.0 / aug[k * ld + k];
    __syncthreads();
    for (int j = tid; j < ld; j += blockDim.x) aug[k * ld + j] *= d;
    __syncthreads();
    for (int e = tid; e < n * ld; e += blockDim.x) {
      const int i = e / ld, j = e - i * ld;
      if (i != k && j != k) aug[i * ld + j] -= aug[i * ld + k] * aug[k * ld + j];
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x)
      if (i != k) aug[i * ld + k] = 0.0;
    __syncthreads();
  }
  // Z_p = A^-1 dC_p
  double* Z = Zs + (size_t)l * npar * n * n;
  for (int e = tid; e < npar * n * n; e += blockDim.x) {
    const int p = e / (n * n), ij = e - p * n * n, i = ij / n, j = ij - i * n;
    const double* D = dC + (((size_t)p * Nl + l) * n) * n;
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += aug[i * ld + n + k] * D[(size_t)k * n + j];
    Z[e] = acc;
  }
  __syncthreads();
  const double lc = 0.5 * (ell[l] + ell[l + 1]);
  const double w = (lc + 0.5) * (ell[l + 1] - ell[l]);
  for (int e = tid; e < npar * npar; e += blockDim.x) {
    const int p = e / npar, q = e - p * npar;
    const double *Zp = Z + (size_t)p * n * n, *Zq = Z + (size_t)q * n * n;
    double acc = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) acc += Zp[i * n + j] * Zq[j * n + i];
    Fl[(size_t)l * npar * npar + e] = w * acc;
  }
}

// chi2[s] = sum over (multipole, block) of w q, one warp per point: lanes stride over the point's contiguous q row
// (coalesced), fixed-order shuffle tree
__global__ void __launch_bounds__(128) photo_chi2_reduce_kernel(int S, int Nl, int nblk, const double* __restrict__ q,
                                                                const double* __restrict__ w, double* __restrict__ chi2) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= S) return;
  const double* qs = q + (size_t)s * Nl * nblk;
  double acc = 0.0;
  for (int e = lane; e < Nl * nblk; e += 32) {
    const int l = e / nblk, b = e - l * nblk;
    const double wl = w[(size_t)b * Nl + l];
    if (wl != 0.0) acc += wl * qs[e];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) chi2[s] = acc;
}

__global__ void photo_add_noise_kernel(int Nl, int Nb, const double* __restrict__ cls0, const double* __restrict__ noise,
                                       double* __restrict__ data) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Nl * Nb * Nb) return;
  const int ij = e % (Nb * Nb);
  data[e] = cls0[e] + ((ij / Nb == ij % Nb) ? noise[ij / Nb] : 0.0);
}

}  // namespace

namespace {
cudaError_t photo_launch_chi2(cfp_ctx* ctx, cudaStream_t st, int S, int Nl, int Nb, int nblk, const int32_t* off_h,
                              const int32_t* n_h, int nmax, const double* theory_d, const double* noise_d,
                              const double* data_d, const double* w_d, double* lb_d, double* lndet_d, double* q_d,
                              double* chi2_d) {
  const ChiBlocks B = chi_blocks(nblk, off_h, n_h);
  const unsigned g_data = (unsigned)(((size_t)Nl + CHR_WARPS - 1) / CHR_WARPS);
  const unsigned g_th = (unsigned)(((size_t)S * Nl + CHR_WARPS - 1) / CHR_WARPS);
#define CHI_LAUNCH(Q)                                                                                                  \
  case Q:                                                                                                              \
    photo_chi2_reg_kernel<true, Q><<<g_data, CHR_WARPS * 32, 0, st>>>(1, Nl, Nb, B, data_d, nullptr, w_d, lb_d, nullptr, \
                                                                      nmax, lndet_d);                                  \
    photo_chi2_reg_kernel<false, Q><<<g_th, CHR_WARPS * 32, 0, st>>>(S, Nl, Nb, B, theory_d, noise_d, w_d, lb_d,        \
                                                                     lndet_d, nmax, q_d);                              \
    break;
  switch ((nmax + 3) / 4) {
    CHI_LAUNCH(1)
    CHI_LAUNCH(2)
    CHI_LAUNCH(3)
    CHI_LAUNCH(4)
    CHI_LAUNCH(5)
    CHI_LAUNCH(6)
    CHI_LAUNCH(7)
    default:
      photo_chi2_reg_kernel<true, 8><<<g_data, CHR_WARPS * 32, 0, st>>>(1, Nl, Nb, B, data_d, nullptr, w_d, lb_d, nullptr, nmax,
                                                                        lndet_d);
      photo_chi2_reg_kernel<false, 8><<<g_th, CHR_WARPS * 32, 0, st>>>(S, Nl, Nb, B, theory_d, noise_d, w_d, lb_d, lndet_d,
                                                                       nmax, q_d);
      break;
  }
#undef CHI_LAUNCH
  photo_chi2_reduce_kernel<<<(S + 3) / 4, 128, 0, st>>>(S, Nl, nblk, q_d, w_d, chi2_d);
  ctx->launches += 3;
  return cudaGetLastError();
}
}  // namespace

struct cfp_photo_plan {
  cfp_ctx* ctx = nullptr;
  const cfp_bank* bank = nullptr;
  int S = 0, NC = 0, Nl = 0, Nz = 0, Nb = 0, npar = 0, nb = 0, ldz = 0, nz4 = 0, ldb = 0, zstride = 0, tab_p = 0;
  int tab_pg = 0;          // table of the clustering rows (= tab_p unless GCph_Tracer = 'clustering')
  double* Tg_d = nullptr;  // its sqrt(P)/chi table (owned only when tab_pg != tab_p)
  int l_begin = 0, l_end = 0;
  double dz = 0.0, kmax = 0.0;
  int *cosmo_of_s_d = nullptr, *kind_d = nullptr;
  double *ell_d = nullptr, *z_d = nullptr, *chi_d = nullptr, *hub_d = nullptr, *wlpref_d = nullptr, *ngal_d = nullptr;
  double *bias_d = nullptr, *ia_d = nullptr, *lmin_d = nullptr, *lmax_d = nullptr, *noise_d = nullptr;
  int NW = 1, NP = 0;     // n(z) window sets; (cosmology, window set) pairs among the stencil points
  int *win_of_s_d = nullptr, *pair_of_s_d = nullptr, *pair_c_d = nullptr, *pair_w_d = nullptr;
  int Kb = 0;             // bias samples per (point, row): Nz, or fewer with a z -> sample map
  bool cls_zeroed = false;  // cls_d was zero-filled (see cfp_photo_plan_set_slice)
  std::vector<void*> chi_tmp;   // device buffers of a launched, not yet collected chi2
  double* chi2_pin = nullptr;   // page-locked copy of its result
  cudaEvent_t chi_up = nullptr;
  bool chi_pending = false;
  int bias_rows = 0;      // rows of the bias array per point: Nb, or 1 when every clustering row shares one vector
  int* zmap_d = nullptr;  // [Nz] or NULL (identity)
  double *sqrtfsky_d = nullptr, *stw_d = nullptr, *stden_d = nullptr, *strden_d = nullptr, *wz_d = nullptr;
  double* stp_w_d = nullptr;
  int *stp_ptr_d = nullptr, *stp_s_d = nullptr;
  double *T_d = nullptr, *eff_d = nullptr, *U_d = nullptr, *V_d = nullptr, *cls_d = nullptr, *Y_d = nullptr;
  double *Fl_d = nullptr, *fisher_d = nullptr;
  bool ran = false;
  cudaStream_t last_st = nullptr;  // stream of the most recent run (the pool releases on the context stream)
  cudaEvent_t kev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // begin / end of this plan's most recent run or chi2 call, on the stream it ran on: fetch waits on run1 (plans of
  // one context may run concurrently on different streams, a context-wide event would be the wrong one)
  cudaEvent_t run0 = nullptr, run1 = nullptr;
  std::vector<void*> owned;
};

static void photo_free(cfp_photo_plan* p) {
  for (void* q : p->chi_tmp) cfp_dev_free(p->ctx, q);
  cfp_pinned_release(p->ctx, p->chi2_pin, (size_t)p->S * sizeof(double));
  if (p->chi_up) cudaEventDestroy(p->chi_up);
  if (p->run0) cudaEventDestroy(p->run0);
  if (p->run1) cudaEventDestroy(p->run1);
  for (auto& e : p->kev)
    if (e) cudaEventDestroy(e);
  for (void* q : p->owned) cfp_dev_free(p->ctx, q);
  delete p;
}

#define PH_UP(T, src, cnt, dst)                                \
  do {                                                         \
    int _rc = cfp_upload<T>(ctx, src, cnt, &(dst));            \
    if (_rc != CFP_OK) {                                       \
      photo_free(pl);                                          \
      return _rc;                                              \
    }                                                          \
    pl->owned.push_back((void*)(dst));                         \
  } while (0)

// buffers every run overwrites completely before reading: allocation only (a likelihood chunk of 4096 points has
// 2.6 GB of them; zero-filling that per plan cost more than the Limber table kernel)
#define PH_ALLOC(T, cnt, dst)                                                              \
  do {                                                                                     \
    (dst) = nullptr;                                                                       \
    if (cfp_dev_alloc(ctx, (void**)&(dst), (size_t)(cnt) * sizeof(T)) != cudaSuccess) {    \
      cfp_set_error("cfp_photo_plan_create: device allocation of %zu bytes failed", (size_t)(cnt) * sizeof(T)); \
      photo_free(pl);                                                                      \
      return CFP_ERR_CUDA;                                                                 \
    }                                                                                      \
    pl->owned.push_back((void*)(dst));                                                     \
  } while (0)

static int photo_plan_create_impl(cfp_ctx* ctx, const cfp_bank* bank, int S, int NC, int Nl, int Nz, int Nb, int npar,
                                  const int32_t* cosmo_of_s_h, const double* ell_h, const double* z_h, double dz,
                                  const double* chi_h, const double* hub_h, const double* wlpref_h,
                                  const int32_t* kind_h, const double* ngal_h, const double* bias_h,
                                  const double* ia_h, const double* lmin_h, const double* lmax_h,
                                  const double* noise_h, const double* sqrtfsky_h, const double* stw_h,
                                  const double* stden_h, double kmax, int tab_p, int Kb, const int32_t* zmap_h,
                                  int NW, const int32_t* win_of_s_h, cfp_photo_plan** out, bool bias_shared = false);

extern "C" int cfp_photo_plan_create(cfp_ctx* ctx, const cfp_bank* bank, int S, int NC, int Nl, int Nz, int Nb,
                                     int npar, const int32_t* cosmo_of_s_h, const double* ell_h, const double* z_h,
                                     double dz, const double* chi_h, const double* hub_h, const double* wlpref_h,
                                     const int32_t* kind_h, const double* ngal_h, const double* bias_h,
                                     const double* ia_h, const double* lmin_h, const double* lmax_h,
                                     const double* noise_h, const double* sqrtfsky_h, const double* stw_h,
                                     const double* stden_h, double kmax, int tab_p, cfp_photo_plan** out) {
  return photo_plan_create_impl(ctx, bank, S, NC, Nl, Nz, Nb, npar, cosmo_of_s_h, ell_h, z_h, dz, chi_h, hub_h, wlpref_h,
                                kind_h, ngal_h, bias_h, ia_h, lmin_h, lmax_h, noise_h, sqrtfsky_h, stw_h, stden_h, kmax,
                                tab_p, Nz, nullptr, 1, nullptr, out);
}

extern "C" int cfp_photo_plan_create_zmap(cfp_ctx* ctx, const cfp_bank* bank, int S, int NC, int Nl, int Nz, int Nb,
                                          int npar, const int32_t* cosmo_of_s_h, const double* ell_h,
                                          const double* z_h, double dz, const double* chi_h, const double* hub_h,
                                          const double* wlpref_h, const int32_t* kind_h, const double* ngal_h,
                                          const double* bias_h, int Kb, const int32_t* zmap_h, const double* ia_h,
                                          const double* lmin_h, const double* lmax_h, const double* noise_h,
                                          const double* sqrtfsky_h, const double* stw_h, const double* stden_h,
                                          double kmax, int tab_p, cfp_photo_plan** out) {
  CFP_REQUIRE(Kb >= 1 && zmap_h != nullptr, "Kb >= 1 and a z map are required");
  for (int i = 0; i < Nz; ++i) CFP_REQUIRE(zmap_h[i] >= 0 && zmap_h[i] < Kb, "z map entry out of range");
  return photo_plan_create_impl(ctx, bank, S, NC, Nl, Nz, Nb, npar, cosmo_of_s_h, ell_h, z_h, dz, chi_h, hub_h, wlpref_h,
                                kind_h, ngal_h, bias_h, ia_h, lmin_h, lmax_h, noise_h, sqrtfsky_h, stw_h, stden_h, kmax,
                                tab_p, Kb, zmap_h, 1, nullptr, out);
}

extern "C" int cfp_photo_plan_create_zmap_shared(cfp_ctx* ctx, const cfp_bank* bank, int S, int NC, int Nl, int Nz,
                                                 int Nb, int npar, const int32_t* cosmo_of_s_h, const double* ell_h,
                                                 const double* z_h, double dz, const double* chi_h, const double* hub_h,
                                                 const double* wlpref_h, const int32_t* kind_h, const double* ngal_h,
                                                 const double* bias_h, int Kb, const int32_t* zmap_h, const double* ia_h,
                                                 const double* lmin_h, const double* lmax_h, const double* noise_h,
                                                 const double* sqrtfsky_h, const double* stw_h, const double* stden_h,
                                                 double kmax, int tab_p, cfp_photo_plan** out) {
  CFP_REQUIRE(Kb >= 1 && zmap_h != nullptr, "Kb >= 1 and a z map are required");
  for (int i = 0; i < Nz; ++i) CFP_REQUIRE(zmap_h[i] >= 0 && zmap_h[i] < Kb, "z map entry out of range");
  return photo_plan_create_impl(ctx, bank, S, NC, Nl, Nz, Nb, npar, cosmo_of_s_h, ell_h, z_h, dz, chi_h, hub_h, wlpref_h,
                                kind_h, ngal_h, bias_h, ia_h, lmin_h, lmax_h, noise_h, sqrtfsky_h, stw_h, stden_h, kmax,
                                tab_p, Kb, zmap_h, 1, nullptr, out, true);
}

extern "C" int cfp_photo_plan_create_ex(cfp_ctx* ctx, const cfp_bank* bank, int S, int NC, int Nl, int Nz, int Nb,
                                        int npar, const int32_t* cosmo_of_s_h, const double* ell_h, const double* z_h,
                                        double dz, const double* chi_h, const double* hub_h, const double* wlpref_h,
                                        const int32_t* kind_h, int NW, const int32_t* win_of_s_h,
                                        const double* ngal_h, const double* bias_h, int Kb, const int32_t* zmap_h,
                                        const double* ia_h, const double* lmin_h, const double* lmax_h,
                                        const double* noise_h, const double* sqrtfsky_h, const double* stw_h,
                                        const double* stden_h, double kmax, int tab_p, cfp_photo_plan** out) {
  CFP_REQUIRE(NW >= 1 && (NW == 1 || win_of_s_h != nullptr), "NW >= 1, and a window index per point when NW > 1");
  if (win_of_s_h)
    for (int s = 0; s < S; ++s) CFP_REQUIRE(win_of_s_h[s] >= 0 && win_of_s_h[s] < NW, "window index out of range");
  if (zmap_h) {
    CFP_REQUIRE(Kb >= 1, "Kb >= 1");
    for (int i = 0; i < Nz; ++i) CFP_REQUIRE(zmap_h[i] >= 0 && zmap_h[i] < Kb, "z map entry out of range");
  }
  return photo_plan_create_impl(ctx, bank, S, NC, Nl, Nz, Nb, npar, cosmo_of_s_h, ell_h, z_h, dz, chi_h, hub_h, wlpref_h,
                                kind_h, ngal_h, bias_h, ia_h, lmin_h, lmax_h, noise_h, sqrtfsky_h, stw_h, stden_h, kmax,
                                tab_p, zmap_h ? Kb : Nz, zmap_h, NW, win_of_s_h, out);
}

static int photo_plan_create_impl(cfp_ctx* ctx, const cfp_bank* bank, int S, int NC, int Nl, int Nz, int Nb, int npar,
                                  const int32_t* cosmo_of_s_h, const double* ell_h, const double* z_h, double dz,
                                  const double* chi_h, const double* hub_h, const double* wlpref_h,
                                  const int32_t* kind_h, const double* ngal_h, const double* bias_h,
                                  const double* ia_h, const double* lmin_h, const double* lmax_h,
                                  const double* noise_h, const double* sqrtfsky_h, const double* stw_h,
                                  const double* stden_h, double kmax, int tab_p, int Kb, const int32_t* zmap_h,
                                  int NW, const int32_t* win_of_s_h, cfp_photo_plan** out, bool bias_shared) {
  CFP_REQUIRE(ctx && bank && out, "NULL ctx/bank/out");
  // (ia_h may be NULL: no intrinsic-alignment window until cfp_photo_plan_set_ia_enla fills it on the device)
  CFP_REQUIRE(cosmo_of_s_h && ell_h && z_h && chi_h && hub_h && wlpref_h && kind_h && ngal_h && bias_h &&
                  lmin_h && lmax_h && noise_h && sqrtfsky_h && stw_h && stden_h,
              "NULL input array");
  CFP_REQUIRE(S >= 1 && NC >= 1 && Nl >= 2 && Nz >= 2 && Nb >= 1 && npar >= 1, "bad dimensions");
  CFP_REQUIRE(Nb <= MAX_PNB * 8, "more than 64 tomographic rows");
  CFP_REQUIRE(npar <= FISH_MAX_PT * 8, "more than 64 parameters");
  CFP_REQUIRE(NC <= bank->n_cosmo, "NC exceeds the bank");
  CFP_REQUIRE(tab_p >= 0 && tab_p < bank->n_tables, "bad table slot");
  for (int s = 0; s < S; ++s) CFP_REQUIRE(cosmo_of_s_h[s] >= 0 && cosmo_of_s_h[s] < NC, "cosmology index out of range");
  for (int c = 0; c < NC; ++c) CFP_REQUIRE(bank->desc(c, tab_p)[0] >= 8, "cosmology lacks the power spectrum table");
  *out = nullptr;
  CFP_CUDA(cudaSetDevice(ctx->device));
  cfp_photo_plan* pl = new cfp_photo_plan();
  pl->ctx = ctx, pl->bank = bank;
  pl->S = S, pl->NC = NC, pl->Nl = Nl, pl->Nz = Nz, pl->Nb = Nb, pl->npar = npar, pl->tab_p = tab_p;
  pl->tab_pg = tab_p;
  pl->nb = (Nb + 7) / 8;
  pl->nz4 = (Nz + 3) / 4 * 4;
  int ldz = pl->nz4;
  while (!((ldz % 16) == 4 || (ldz % 16) == 12)) ldz += 4;  // conflict-free DMMA fragment loads
  pl->ldz = ldz;
  if ((size_t)pl->nb * 8 * ldz * sizeof(double) > 220 * 1024) {
    cfp_set_error("cfp_photo_plan_create: Nb_pad x Nz = %d x %d does not fit the shared-memory tile of the C_ell kernel "
                  "(limit Nb_pad*Nz*8 B <= 220 KB)", pl->nb * 8, ldz);
    photo_free(pl);
    return CFP_ERR_INVALID;
  }
  // row stride of the Nb x Nb tiles and stride between the Z_p blocks: multiples of 4 that are 4 or 12 mod 16,
  // so that the m8n8k4 fragment loads of photo_fisher_kernel hit 16 distinct 8-byte banks per half warp
  pl->ldb = Nb;
  while (pl->ldb % 16 != 4 && pl->ldb % 16 != 12) ++pl->ldb;
  pl->zstride = Nb * pl->ldb;
  while (pl->zstride % 16 != 4 && pl->zstride % 16 != 12) pl->zstride += 4;
  pl->dz = dz, pl->kmax = kmax;
  pl->l_begin = 0, pl->l_end = Nl - 1;
  std::vector<double> wz(ldz, 0.0);
  for (int i = 0; i < Nz; ++i) wz[i] = dz * ((i == 0 || i == Nz - 1) ? 0.5 : 1.0);  // photo_obs.py:921
  PH_UP(int, cosmo_of_s_h, (size_t)S, pl->cosmo_of_s_d);
  PH_UP(int, kind_h, (size_t)Nb, pl->kind_d);
  PH_UP(double, ell_h, (size_t)Nl, pl->ell_d);
  PH_UP(double, z_h, (size_t)Nz, pl->z_d);
  PH_UP(double, chi_h, (size_t)NC * Nz, pl->chi_d);
  PH_UP(double, hub_h, (size_t)NC * Nz, pl->hub_d);
  PH_UP(double, wlpref_h, (size_t)NC, pl->wlpref_d);
  PH_UP(double, ngal_h, (size_t)NW * Nb * Nz, pl->ngal_d);
  pl->NW = NW;
  {
    // distinct (cosmology, window set) pairs among the stencil points: one lensing-efficiency table each
    std::vector<int> pair_of_s(S), pair_c, pair_w;
    if (NW > 1) {  // (with one window set the pair index is the cosmology index: nothing to search)
      std::unordered_map<uint64_t, int> seen;  // a batch may bring thousands of cosmologies: no linear search
      seen.reserve((size_t)S * 2);
      for (int s = 0; s < S; ++s) {
        const int c = cosmo_of_s_h[s], w = win_of_s_h ? win_of_s_h[s] : 0;
        const uint64_t key = ((uint64_t)(uint32_t)c << 32) | (uint32_t)w;
        auto it = seen.find(key);
        if (it == seen.end()) {
          it = seen.emplace(key, (int)pair_c.size()).first;
          pair_c.push_back(c), pair_w.push_back(w);
        }
        pair_of_s[s] = it->second;
      }
    }
    if (NW == 1) {  // keep the plain layout: pair index = cosmology index, every cosmology has a table
      pl->NP = NC;
    } else {
      pl->NP = (int)pair_c.size();
      PH_UP(int, win_of_s_h, (size_t)S, pl->win_of_s_d);
      PH_UP(int, pair_of_s.data(), (size_t)S, pl->pair_of_s_d);
      PH_UP(int, pair_c.data(), pair_c.size(), pl->pair_c_d);
      PH_UP(int, pair_w.data(), pair_w.size(), pl->pair_w_d);
    }
  }
  pl->Kb = Kb;
  pl->bias_rows = bias_shared ? 1 : Nb;  // shared: one bias vector per point for every clustering row
  PH_UP(double, bias_h, (size_t)S * pl->bias_rows * Kb, pl->bias_d);
  if (zmap_h) PH_UP(int, zmap_h, (size_t)Nz, pl->zmap_d);
  PH_UP(double, ia_h, (size_t)S * Nz, pl->ia_d);
  PH_UP(double, lmin_h, (size_t)Nb, pl->lmin_d);
  PH_UP(double, lmax_h, (size_t)Nb, pl->lmax_d);
  PH_UP(double, noise_h, (size_t)Nb, pl->noise_d);
  PH_UP(double, sqrtfsky_h, (size_t)Nb, pl->sqrtfsky_d);
  PH_UP(double, stw_h, (size_t)npar * S, pl->stw_d);
  PH_UP(double, stden_h, (size_t)npar, pl->stden_d);
  std::vector<double> strden(npar);
  for (int p = 0; p < npar; ++p) strden[p] = 1.0 / stden_h[p];
  PH_UP(double, strden.data(), (size_t)npar, pl->strden_d);
  PH_UP(double, wz.data(), (size_t)ldz, pl->wz_d);
  {
    std::vector<int> ptr(npar + 1, 0), ss;
    std::vector<double> ww;
    for (int p = 0; p < npar; ++p) {
      for (int q = 0; q < S; ++q)
        if (stw_h[(size_t)p * S + q] != 0.0) {
          ss.push_back(q);
          ww.push_back(stw_h[(size_t)p * S + q]);
        }
      ptr[p + 1] = (int)ss.size();
    }
    if (ss.empty()) {
      ss.push_back(0);
      ww.push_back(0.0);
    }
    PH_UP(int, ptr.data(), ptr.size(), pl->stp_ptr_d);
    PH_UP(int, ss.data(), ss.size(), pl->stp_s_d);
    PH_UP(double, ww.data(), ww.size(), pl->stp_w_d);
  }
  PH_UP(double, nullptr, (size_t)NC * Nl * Nz, pl->T_d);
  PH_UP(double, nullptr, (size_t)pl->NP * Nb * Nz, pl->eff_d);
  PH_ALLOC(double, (size_t)S * pl->nb * 8 * ldz, pl->U_d);  // (the window kernel writes every element, padding included)
  PH_ALLOC(double, (size_t)S * pl->nb * 8 * ldz, pl->V_d);
  if ((size_t)S * Nl * Nb * Nb * sizeof(double) > ((size_t)64 << 20)) {
    // a likelihood batch: every multipole is written by every run (a multipole SLICE leaves the rest untouched,
    // which only Fisher plans -- a few MB, zero-filled below -- are asked for)
    PH_ALLOC(double, (size_t)S * Nl * Nb * Nb, pl->cls_d);
  } else {
    PH_UP(double, nullptr, (size_t)S * Nl * Nb * Nb, pl->cls_d);
    pl->cls_zeroed = true;
  }
  PH_UP(double, nullptr, (size_t)(Nl - 1) * npar * pl->zstride, pl->Y_d);
  PH_UP(double, nullptr, (size_t)(Nl - 1) * npar * npar, pl->Fl_d);
  PH_UP(double, nullptr, (size_t)npar * npar, pl->fisher_d);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    cfp_set_error("cfp_photo_plan_create: upload failed");
    photo_free(pl);
    return CFP_ERR_CUDA;
  }
  *out = pl;
  return CFP_OK;
}

extern "C" int cfp_photo_plan_destroy(cfp_photo_plan* plan) {
  if (!plan) return CFP_OK;
  cudaSetDevice(plan->ctx->device);
  // the buffers return to the pool on the context stream: wait for THIS plan's work on the caller's stream -- its own
  // end-of-run events, not the whole stream, which in a pipeline of batches already holds the next plan's kernels
  if (plan->last_st && plan->last_st != plan->ctx->stream) {
    if (plan->chi_pending && plan->chi_up) {
      cudaEventSynchronize(plan->chi_up);
    } else if (plan->run1) {
      cudaEventSynchronize(plan->run1);
    } else {
      cudaStreamSynchronize(plan->last_st);
    }
  }
  photo_free(plan);
  return CFP_OK;
}

// eNLA intrinsic-alignment window of every point on the job's z grid, from the point's (amplitude factor, eta, beta):
// W_IA(z_j) = fac * ((1 + z_j) / (1 + z_piv))^eta * (L(z_j)^beta / D(z_j)) on the nuisance grid (nuisance.py:190-230),
// then the degree-1 interpolation to the z samples (the products in the order numpy forms them on the host path).
namespace {
__global__ void photo_ia_kernel(int S, int Nz, int nia, const int* __restrict__ cosmo_of_s, const double* __restrict__ zdep,
                                const double* __restrict__ lum, const double* __restrict__ growth,
                                const double* __restrict__ fac, const double* __restrict__ eta,
                                const double* __restrict__ beta, const int* __restrict__ jj, const double* __restrict__ tt,
                                double* __restrict__ ia) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)S * Nz) return;
  const int s = (int)(idx / Nz), iz = (int)(idx - (size_t)s * Nz);
  const int j = jj[iz];
  const double t = tt[iz];
  const double* g = growth + (size_t)cosmo_of_s[s] * nia;
  const double f = fac[s], e = eta[s], b = beta[s];
  const double w0 = __dmul_rn(__dmul_rn(f, pow(zdep[j], e)), __ddiv_rn(pow(lum[j], b), g[j]));
  const double w1 = __dmul_rn(__dmul_rn(f, pow(zdep[j + 1], e)), __ddiv_rn(pow(lum[j + 1], b), g[j + 1]));
  ia[idx] = __dadd_rn(__dmul_rn(w0, 1.0 - t), __dmul_rn(w1, t));
}
}  // namespace

extern "C" int cfp_photo_plan_set_ia_enla(cfp_photo_plan* plan, int nia, const double* zdep_h, const double* lum_h,
                                          const double* growth_h, const double* fac_h, const double* eta_h,
                                          const double* beta_h, const int32_t* j_h, const double* t_h) {
  CFP_REQUIRE(plan != nullptr, "plan is NULL");
  CFP_REQUIRE(zdep_h && lum_h && growth_h && fac_h && eta_h && beta_h && j_h && t_h, "NULL input array");
  CFP_REQUIRE(nia >= 2, "the nuisance redshift grid needs two samples");
  for (int i = 0; i < plan->Nz; ++i) CFP_REQUIRE(j_h[i] >= 0 && j_h[i] <= nia - 2, "interpolation index out of range");
  cfp_ctx* ctx = plan->ctx;
  cfp_photo_plan* pl = plan;
  CFP_CUDA(cudaSetDevice(ctx->device));
  double *zd = nullptr, *lu = nullptr, *gr = nullptr, *fa = nullptr, *et = nullptr, *be = nullptr, *td = nullptr;
  int* jd = nullptr;
#define IA_UP(T, src, cnt, dst)                                   \
  do {                                                            \
    int _rc = cfp_upload<T>(ctx, src, cnt, &(dst));               \
    if (_rc != CFP_OK) return _rc;                                \
    pl->owned.push_back((void*)(dst));                            \
  } while (0)
  IA_UP(double, zdep_h, (size_t)nia, zd);
  IA_UP(double, lum_h, (size_t)nia, lu);
  IA_UP(double, growth_h, (size_t)pl->NC * nia, gr);
  IA_UP(double, fac_h, (size_t)pl->S, fa);
  IA_UP(double, eta_h, (size_t)pl->S, et);
  IA_UP(double, beta_h, (size_t)pl->S, be);
  IA_UP(int, j_h, (size_t)pl->Nz, jd);
  IA_UP(double, t_h, (size_t)pl->Nz, td);
#undef IA_UP
  const size_t tot = (size_t)pl->S * pl->Nz;
  photo_ia_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(pl->S, pl->Nz, nia, pl->cosmo_of_s_d, zd, lu, gr, fa, et, be,
                                                                         jd, td, pl->ia_d);
  ctx->launches++;
  CFP_CUDA(cudaGetLastError());
  return CFP_OK;  // (pageable host arrays are staged before cudaMemcpyAsync returns, as in plan creation)
}

extern "C" int cfp_photo_plan_set_gc_table(cfp_photo_plan* plan, int tab_p_gc) {
  CFP_REQUIRE(plan != nullptr, "plan is NULL");
  CFP_REQUIRE(tab_p_gc >= 0 && tab_p_gc < plan->bank->n_tables, "bad table slot");
  for (int c = 0; c < plan->NC; ++c)
    CFP_REQUIRE(plan->bank->desc(c, tab_p_gc)[0] >= 8, "cosmology lacks the clustering power spectrum table");
  cfp_ctx* ctx = plan->ctx;
  CFP_CUDA(cudaSetDevice(ctx->device));
  if (tab_p_gc != plan->tab_p && !plan->Tg_d) {
    CFP_CUDA(cfp_dev_alloc(ctx, (void**)&plan->Tg_d, (size_t)plan->NC * plan->Nl * plan->Nz * sizeof(double)));
    plan->owned.push_back((void*)plan->Tg_d);
    CFP_CUDA(cudaMemsetAsync(plan->Tg_d, 0, (size_t)plan->NC * plan->Nl * plan->Nz * sizeof(double), ctx->stream));
    CFP_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  plan->tab_pg = tab_p_gc;
  return CFP_OK;
}

extern "C" int cfp_photo_plan_set_slice(cfp_photo_plan* plan, int l_begin, int l_end) {
  CFP_REQUIRE(plan != nullptr, "plan is NULL");
  CFP_REQUIRE(l_begin >= 0 && l_begin <= l_end && l_end <= plan->Nl - 1, "slice out of range");
  plan->l_begin = l_begin;
  plan->l_end = l_end;
  if (!plan->cls_zeroed && !(l_begin == 0 && l_end == plan->Nl - 1)) {
    // a partial run leaves the spectra of the other multipoles untouched: define them (zero) once
    CFP_CUDA(cudaSetDevice(plan->ctx->device));
    CFP_CUDA(cudaMemsetAsync(plan->cls_d, 0, (size_t)plan->S * plan->Nl * plan->Nb * plan->Nb * sizeof(double),
                             plan->ctx->stream));
    CFP_CUDA(cudaStreamSynchronize(plan->ctx->stream));
    plan->cls_zeroed = true;
  }
  return CFP_OK;
}

template <int NB, bool TWO, int WARPS>
static cudaError_t launch_cls_w(const ClsParams& P, dim3 grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(photo_cls_kernel<NB, TWO, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  photo_cls_kernel<NB, TWO, WARPS><<<grid, WARPS * 32, smem, st>>>(P);
  return cudaGetLastError();
}

template <int NB>
static cudaError_t launch_cls(const ClsParams& P, dim3 grid, size_t smem, cudaStream_t st, int threads) {
  const bool two = P.Tg != P.T;
  switch (threads / 32) {
    case 8: return two ? launch_cls_w<NB, true, 8>(P, grid, smem, st) : launch_cls_w<NB, false, 8>(P, grid, smem, st);
    case 16: return two ? launch_cls_w<NB, true, 16>(P, grid, smem, st) : launch_cls_w<NB, false, 16>(P, grid, smem, st);
    default: return two ? launch_cls_w<NB, true, 4>(P, grid, smem, st) : launch_cls_w<NB, false, 4>(P, grid, smem, st);
  }
}

static int photo_run_cls(cfp_photo_plan* pl, cudaStream_t st, bool all_ell);
static int photo_run_fisher(cfp_photo_plan* pl, cudaStream_t st);

extern "C" int cfp_photo_plan_run(cfp_photo_plan* pl, void* stream) {
  CFP_REQUIRE(pl != nullptr, "plan is NULL");
  cfp_ctx* ctx = pl->ctx;
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  if (!pl->run0) {
    CFP_CUDA(cudaEventCreate(&pl->run0));
    CFP_CUDA(cudaEventCreate(&pl->run1));
  }
  CFP_CUDA(cudaEventRecord(pl->run0, st));
  int rc = photo_run_cls(pl, st, false);
  if (rc != CFP_OK) return rc;
  rc = photo_run_fisher(pl, st);
  if (rc != CFP_OK) return rc;
  CFP_CUDA(cudaEventRecord(pl->run1, st));
  pl->ran = true;
  return CFP_OK;
}

static int photo_run_cls(cfp_photo_plan* pl, cudaStream_t st, bool all_ell) {
  cfp_ctx* ctx = pl->ctx;
  pl->last_st = st;
  for (auto& e : pl->kev)
    if (!e) CFP_CUDA(cudaEventCreate(&e));
  CFP_CUDA(cudaEventRecord(pl->kev[0], st));
  {
    // the multipoles of this run (a slice of a sharded forecast computes its own rows of the Limber table only; the
    // table is zero-filled at plan creation, so the rows of other ranks read as zeros)
    const int l0 = all_ell ? 0 : pl->l_begin;
    const int nl = (all_ell ? pl->Nl : std::min(pl->l_end + 1, pl->Nl)) - l0;
    const size_t tot = (size_t)pl->NC * nl * pl->Nz;
    if (tot > 0) {
      photo_sqrtp_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(
          pl->bank->blob_d, pl->bank->desc_d, pl->bank->n_tables, pl->tab_p, pl->NC, pl->Nl, pl->Nz, l0, nl, pl->ell_d,
          pl->z_d, pl->chi_d, pl->kmax, pl->T_d);
      ctx->launches++;
      CFP_CUDA(cudaGetLastError());
      if (pl->tab_pg != pl->tab_p) {
        photo_sqrtp_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(
            pl->bank->blob_d, pl->bank->desc_d, pl->bank->n_tables, pl->tab_pg, pl->NC, pl->Nl, pl->Nz, l0, nl, pl->ell_d,
            pl->z_d, pl->chi_d, pl->kmax, pl->Tg_d);
        ctx->launches++;
        CFP_CUDA(cudaGetLastError());
      }
    }
  }
  CFP_CUDA(cudaEventRecord(pl->kev[1], st));
  {
    const size_t esm = (size_t)(3 * pl->Nb + 1) * pl->Nz * sizeof(double);
    if (esm <= 200 * 1024) {
      if (esm > 48 * 1024)
        CFP_CUDA(cudaFuncSetAttribute(photo_eff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm));
      photo_eff_kernel<<<pl->NP, 256, esm, st>>>(pl->NP, pl->Nb, pl->Nz, pl->kind_d, pl->pair_c_d, pl->pair_w_d,
                                                pl->ngal_d, pl->chi_d, pl->dz, pl->eff_d);
    } else {
      const int tot = pl->NP * pl->Nb;
      photo_eff_kernel_global<<<(tot + 63) / 64, 64, 0, st>>>(pl->NP, pl->Nb, pl->Nz, pl->kind_d, pl->pair_c_d,
                                                             pl->pair_w_d, pl->ngal_d, pl->chi_d, pl->dz, pl->eff_d);
    }
    ctx->launches++;
    CFP_CUDA(cudaGetLastError());
  }
  {
    const size_t tot = (size_t)pl->S * pl->nb * 8 * pl->ldz;
    photo_window_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
        pl->S, pl->Nb, pl->Nz, pl->ldz, pl->nb * 8, pl->cosmo_of_s_d, pl->win_of_s_d, pl->pair_of_s_d, pl->kind_d,
        pl->z_d, pl->chi_d, pl->hub_d,
        pl->wlpref_d, pl->ngal_d, pl->bias_d, pl->Kb, pl->bias_rows, pl->zmap_d, pl->ia_d, pl->eff_d, pl->U_d, pl->V_d);
    ctx->launches++;
    CFP_CUDA(cudaGetLastError());
  }
  // C_ell of every stencil point at every multipole of the slice (+ the right edge of the last bin is
  // not needed: the Fisher uses left-edge values only, cosmicfish.py:704-723) -- but C_ell is part of
  // the public result, so all multipoles [l_begin, l_end] are produced.
  {
    ClsParams P;
    P.S = pl->S, P.Nl = pl->Nl, P.Nz = pl->Nz, P.Nb = pl->Nb, P.nb = pl->nb, P.ldz = pl->ldz, P.nz4 = pl->nz4;
    P.l_begin = all_ell ? 0 : pl->l_begin;
    P.l_end = all_ell ? pl->Nl : std::min(pl->l_end + 1, pl->Nl);
    const int nl = P.l_end - P.l_begin;
    const size_t smem = (size_t)pl->nb * 8 * pl->ldz * sizeof(double);
    // multipoles per CTA: one wave of as many CTAs as fit beside their staged windows (at most 8 per SM)
    const int per_sm = (int)std::min<size_t>(8, std::max<size_t>(1, (size_t)220 * 1024 / std::max<size_t>(smem, 1)));
    const int slots = per_sm * ctx->sm_count;
    // warps per CTA: the CTAs of a point share nothing but re-stage its windows (52 KB for 26 rows), so more warps
    // beside one staged copy beat more CTAs; CFP_CLS_WARPS overrides for A/B timing
    static const int cls_warps_env = []() {
      const char* e = getenv("CFP_CLS_WARPS");
      const int v = e ? atoi(e) : 0;
      return (v == 4 || v == 8 || v == 16) ? v : 0;
    }();
    const int cls_warps = cls_warps_env ? cls_warps_env : CLS_WARPS;
    int lch = std::max(cls_warps, (nl * pl->S + slots - 1) / slots);
    lch = (lch + cls_warps - 1) / cls_warps * cls_warps;
    {
      const char* e = getenv("CFP_CLS_LCH");  // (A/B: multipoles per CTA)
      if (e && atoi(e) > 0) lch = atoi(e);
    }
    P.lch = lch;
    P.cosmo_of_s = pl->cosmo_of_s_d, P.U = pl->U_d, P.V = pl->V_d, P.T = pl->T_d, P.wz = pl->wz_d;
    P.Tg = pl->tab_pg != pl->tab_p ? pl->Tg_d : pl->T_d, P.kind = pl->kind_d;
    P.ell = pl->ell_d, P.lmin = pl->lmin_d, P.lmax = pl->lmax_d, P.cls = pl->cls_d;
    CFP_CUDA(cudaEventRecord(pl->kev[2], st));
    const dim3 grid((nl + lch - 1) / lch, pl->S);
    cudaError_t e = cudaSuccess;
    if (nl > 0) {
      switch (pl->nb) {
        case 1: e = launch_cls<1>(P, grid, smem, st, cls_warps * 32); break;
        case 2: e = launch_cls<2>(P, grid, smem, st, cls_warps * 32); break;
        case 3: e = launch_cls<3>(P, grid, smem, st, cls_warps * 32); break;
        case 4: e = launch_cls<4>(P, grid, smem, st, cls_warps * 32); break;
        case 5: e = launch_cls<5>(P, grid, smem, st, cls_warps * 32); break;
        case 6: e = launch_cls<6>(P, grid, smem, st, cls_warps * 32); break;
        case 7: e = launch_cls<7>(P, grid, smem, st, cls_warps * 32); break;
        default: e = launch_cls<8>(P, grid, smem, st, cls_warps * 32); break;
      }
      ctx->launches++;
    }
    if (e != cudaSuccess) {
      cfp_set_error("photo_cls_kernel launch: %s (smem %zu B)", cudaGetErrorString(e), smem);
      return CFP_ERR_CUDA;
    }
    CFP_CUDA(cudaEventRecord(pl->kev[3], st));
  }
  return CFP_OK;
}

static int photo_run_fisher(cfp_photo_plan* pl, cudaStream_t st) {
  cfp_ctx* ctx = pl->ctx;
  {
    FishParams P;
    P.S = pl->S, P.Nl = pl->Nl, P.Nb = pl->Nb, P.npar = pl->npar, P.l_begin = pl->l_begin, P.l_end = pl->l_end;
    P.ldb = pl->ldb, P.zstride = pl->zstride;
    P.cls = pl->cls_d, P.ell = pl->ell_d, P.noise = pl->noise_d, P.sqrtfsky = pl->sqrtfsky_d;
    P.stp_ptr = pl->stp_ptr_d, P.stp_s = pl->stp_s_d, P.stp_w = pl->stp_w_d;
    P.stden = pl->stden_d, P.strden = pl->strden_d, P.Y = pl->Y_d, P.Fl = pl->Fl_d;
    const int nbins = pl->l_end - pl->l_begin;
    const int PT = (pl->npar + 7) / 8;
    size_t smem = ((size_t)pl->Nb * pl->ldb + (size_t)FISH_WARPS * PT * 64) * sizeof(double);
    const size_t zsm = (size_t)pl->npar * pl->zstride * sizeof(double);
    const bool zs = smem + zsm <= 220 * 1024;
    if (zs) smem += zsm;
    void (*kern)(FishParams) = nullptr;
    if (pl->Nb <= 16)
      kern = zs ? photo_fisher_kernel<4, true> : photo_fisher_kernel<4, false>;
    else if (pl->Nb <= 32)
      kern = zs ? photo_fisher_kernel<8, true> : photo_fisher_kernel<8, false>;
    else
      kern = zs ? photo_fisher_kernel<16, true> : photo_fisher_kernel<16, false>;
    CFP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CFP_CUDA(cudaEventRecord(pl->kev[4], st));
    if (nbins > 0) {
      kern<<<nbins, FISH_THREADS, smem, st>>>(P);
      ctx->launches++;
      CFP_CUDA(cudaGetLastError());
    }
    CFP_CUDA(cudaEventRecord(pl->kev[5], st));
    const int n = pl->npar * pl->npar;
    photo_reduce_kernel<<<(n + 7) / 8, 256, 0, st>>>(pl->Fl_d, nbins, n, pl->fisher_d);
    ctx->launches++;
    CFP_CUDA(cudaGetLastError());
  }
  return CFP_OK;
}

// chi^2 of every point of the plan, in two halves so that a caller can keep the device busy while it prepares the next
// batch on the host: _launch uploads the block description, enqueues C_ell, the factorisations and the multipole sums
// on `stream` (NULL: the context stream) and a copy of the result into a page-locked buffer of the plan, and RETURNS
// without waiting; _collect waits for that copy and hands out chi2_h[S].  cfp_photo_plan_chi2 = both.
static void photo_chi_release(cfp_photo_plan* pl) {
  for (void* q : pl->chi_tmp) cfp_dev_free(pl->ctx, q);
  pl->chi_tmp.clear();
  pl->chi_pending = false;
}

extern "C" int cfp_photo_plan_chi2_launch(cfp_photo_plan* pl, const double* data_h, int nblk, const int32_t* blk_off_h,
                                          const int32_t* blk_n_h, const double* blk_w_h, void* stream) {
  CFP_REQUIRE(pl && blk_off_h && blk_n_h && blk_w_h, "NULL argument");
  CFP_REQUIRE(nblk >= 1 && nblk <= 8, "1..8 blocks");
  CFP_REQUIRE(!pl->chi_pending, "a chi2 launch of this plan has not been collected yet");
  for (int b = 0; b < nblk; ++b)
    CFP_REQUIRE(blk_n_h[b] >= 1 && blk_n_h[b] <= 32 && blk_off_h[b] >= 0 && blk_off_h[b] + blk_n_h[b] <= pl->Nb,
                "block out of range (at most 32 rows per block)");
  cfp_ctx* ctx = pl->ctx;
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const int S = pl->S, Nl = pl->Nl, Nb = pl->Nb;
  int nmax = 1;
  for (int b = 0; b < nblk; ++b) nmax = std::max(nmax, (int)blk_n_h[b]);
  int *off_d = nullptr, *n_d = nullptr;
  double *w_d = nullptr, *data_d = nullptr, *lndet_d = nullptr, *q_d = nullptr, *chi2_d = nullptr, *zero_d = nullptr,
         *lb_d = nullptr;
  int rc = CFP_OK;
#define CHI_TRY(x, ptr)                      \
  do {                                       \
    rc = (x);                                \
    if (rc != CFP_OK) {                      \
      photo_chi_release(pl);                 \
      return rc;                             \
    }                                        \
    pl->chi_tmp.push_back((void*)(ptr));     \
  } while (0)
  CHI_TRY(cfp_upload(ctx, blk_off_h, (size_t)nblk, &off_d), off_d);
  CHI_TRY(cfp_upload(ctx, blk_n_h, (size_t)nblk, &n_d), n_d);
  CHI_TRY(cfp_upload(ctx, blk_w_h, (size_t)nblk * Nl, &w_d), w_d);
  CHI_TRY(cfp_upload<double>(ctx, data_h, (size_t)Nl * Nb * Nb, &data_d), data_d);
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)Nl * nblk, &lndet_d), lndet_d);
  {  // q is written for every (point, multipole, block): allocation only
    cudaError_t ea = cfp_dev_alloc(ctx, (void**)&q_d, (size_t)S * Nl * nblk * sizeof(double));
    CHI_TRY(ea == cudaSuccess ? CFP_OK : CFP_ERR_CUDA, q_d);
  }
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)S, &chi2_d), chi2_d);
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)Nb, &zero_d), zero_d);
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)Nl * nblk * nmax * nmax, &lb_d), lb_d);
#undef CHI_TRY
  if (!pl->chi2_pin) pl->chi2_pin = (double*)cfp_pinned_acquire(ctx, (size_t)S * sizeof(double));
  if (!pl->chi2_pin) {
    photo_chi_release(pl);
    cfp_set_error("cfp_photo_plan_chi2: cannot allocate the page-locked result buffer");
    return CFP_ERR_CUDA;
  }
  if (!pl->run0) {
    cudaEventCreate(&pl->run0);
    cudaEventCreate(&pl->run1);
  }
  if (!pl->chi_up) cudaEventCreate(&pl->chi_up);
  if (st != ctx->stream) {  // the uploads ran on the context stream: order the work stream behind them (no host wait)
    cudaEventRecord(pl->chi_up, ctx->stream);
    cudaStreamWaitEvent(st, pl->chi_up, 0);
  }
  cudaEventRecord(pl->run0, st);
  rc = photo_run_cls(pl, st, true);
  if (rc != CFP_OK) {
    photo_chi_release(pl);
    return rc;
  }
  if (!data_h) {
    photo_add_noise_kernel<<<(Nl * Nb * Nb + 255) / 256, 256, 0, st>>>(Nl, Nb, pl->cls_d, pl->noise_d, data_d);
    ctx->launches++;
  }
  // ln det and Cholesky factor of the data blocks, then every (point, multipole), then the multipole sums
  cudaError_t e = photo_launch_chi2(ctx, st, S, Nl, Nb, nblk, blk_off_h, blk_n_h, nmax, pl->cls_d, pl->noise_d, data_d, w_d,
                                    lb_d, lndet_d, q_d, chi2_d);
  cudaEventRecord(pl->run1, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(pl->chi2_pin, chi2_d, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaEventRecord(pl->chi_up, st);  // (re-used: marks the end of the result copy)
  if (e != cudaSuccess) {
    cudaStreamSynchronize(st);
    photo_chi_release(pl);
    cfp_set_error("cfp_photo_plan_chi2: %s", cudaGetErrorString(e));
    return CFP_ERR_CUDA;
  }
  pl->last_st = st;
  pl->chi_pending = true;
  return CFP_OK;
}

extern "C" int cfp_photo_plan_chi2_collect(cfp_photo_plan* pl, double* chi2_h) {
  CFP_REQUIRE(pl && chi2_h, "NULL argument");
  CFP_REQUIRE(pl->chi_pending, "no chi2 launch to collect");
  cfp_ctx* ctx = pl->ctx;
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaError_t e = cudaEventSynchronize(pl->chi_up);
  float ms = 0.f;
  if (e == cudaSuccess && cudaEventElapsedTime(&ms, pl->run0, pl->run1) == cudaSuccess) ctx->last_ms = ms;
  if (e == cudaSuccess) std::memcpy(chi2_h, pl->chi2_pin, (size_t)pl->S * sizeof(double));
  photo_chi_release(pl);
  if (e != cudaSuccess) {
    cfp_set_error("cfp_photo_plan_chi2: %s", cudaGetErrorString(e));
    return CFP_ERR_CUDA;
  }
  pl->ran = true;
  return CFP_OK;
}

extern "C" int cfp_photo_plan_chi2(cfp_photo_plan* pl, const double* data_h, int nblk, const int32_t* blk_off_h,
                                   const int32_t* blk_n_h, const double* blk_w_h, double* chi2_h, void* stream) {
  CFP_REQUIRE(chi2_h != nullptr, "NULL argument");
  const int rc = cfp_photo_plan_chi2_launch(pl, data_h, nblk, blk_off_h, blk_n_h, blk_w_h, stream);
  return rc == CFP_OK ? cfp_photo_plan_chi2_collect(pl, chi2_h) : rc;
}

extern "C" double cfp_photo_plan_kernel_ms(cfp_photo_plan* pl, int which) {
  if (!pl || !pl->ran || which < 0 || which > 2) return -1.0;
  cudaSetDevice(pl->ctx->device);
  if (cudaEventSynchronize(pl->kev[5]) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, pl->kev[2 * which], pl->kev[2 * which + 1]) != cudaSuccess) return -1.0;
  return ms;
}

extern "C" const double* cfp_photo_plan_fisher_d(const cfp_photo_plan* plan) { return plan ? plan->fisher_d : nullptr; }

extern "C" int cfp_photo_plan_fetch(cfp_photo_plan* pl, double* fisher_h, double* cls_h, double* sqrtp_h) {
  CFP_REQUIRE(pl != nullptr, "plan is NULL");
  if (!pl->ran) {
    cfp_set_error("cfp_photo_plan_fetch: plan has not been run");
    return CFP_ERR_STATE;
  }
  cfp_ctx* ctx = pl->ctx;
  CFP_CUDA(cudaSetDevice(ctx->device));
  CFP_CUDA(cudaEventSynchronize(pl->run1));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, pl->run0, pl->run1) == cudaSuccess) ctx->last_ms = ms;
  if (fisher_h)
    CFP_CUDA(cudaMemcpy(fisher_h, pl->fisher_d, (size_t)pl->npar * pl->npar * sizeof(double), cudaMemcpyDeviceToHost));
  if (cls_h)
    CFP_CUDA(cudaMemcpy(cls_h, pl->cls_d, (size_t)pl->S * pl->Nl * pl->Nb * pl->Nb * sizeof(double), cudaMemcpyDeviceToHost));
  if (sqrtp_h)
    CFP_CUDA(cudaMemcpy(sqrtp_h, pl->T_d, (size_t)pl->NC * pl->Nl * pl->Nz * sizeof(double), cudaMemcpyDeviceToHost));
  return CFP_OK;
}

extern "C" int cfp_cells_chi2(cfp_ctx* ctx, int S, int Nl, int Nb, const double* theory_h, const double* data_h, int nblk,
                              const int32_t* blk_off_h, const int32_t* blk_n_h, const double* blk_w_h, double* chi2_h) {
  CFP_REQUIRE(ctx && theory_h && data_h && blk_off_h && blk_n_h && blk_w_h && chi2_h, "NULL argument");
  CFP_REQUIRE(S >= 1 && Nl >= 1 && Nb >= 1 && nblk >= 1 && nblk <= 8, "bad dimensions");
  int nmax = 1;
  for (int b = 0; b < nblk; ++b) {
    CFP_REQUIRE(blk_n_h[b] >= 1 && blk_n_h[b] <= 32 && blk_off_h[b] >= 0 && blk_off_h[b] + blk_n_h[b] <= Nb,
                "block out of range (at most 32 rows per block)");
    nmax = std::max(nmax, (int)blk_n_h[b]);
  }
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int *off_d = nullptr, *n_d = nullptr;
  double *w_d = nullptr, *th_d = nullptr, *data_d = nullptr, *zero_d = nullptr, *lndet_d = nullptr, *q_d = nullptr, *chi2_d = nullptr;
  double* lb_d = nullptr;
  auto cleanup = [&]() {
    for (void* q : {(void*)off_d, (void*)n_d, (void*)w_d, (void*)th_d, (void*)data_d, (void*)zero_d, (void*)lndet_d,
                    (void*)q_d, (void*)chi2_d, (void*)lb_d})
      cfp_dev_free(ctx, q);
  };
  int rc;
#define CHI_TRY(x)      \
  do {                  \
    rc = (x);           \
    if (rc != CFP_OK) { \
      cleanup();        \
      return rc;        \
    }                   \
  } while (0)
  CHI_TRY(cfp_upload(ctx, blk_off_h, (size_t)nblk, &off_d));
  CHI_TRY(cfp_upload(ctx, blk_n_h, (size_t)nblk, &n_d));
  CHI_TRY(cfp_upload(ctx, blk_w_h, (size_t)nblk * Nl, &w_d));
  CHI_TRY(cfp_upload(ctx, theory_h, (size_t)S * Nl * Nb * Nb, &th_d));
  CHI_TRY(cfp_upload(ctx, data_h, (size_t)Nl * Nb * Nb, &data_d));
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)Nb, &zero_d));
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)Nl * nblk, &lndet_d));
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)S * Nl * nblk, &q_d));
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)S, &chi2_d));
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)Nl * nblk * nmax * nmax, &lb_d));
#undef CHI_TRY
  cudaError_t e = photo_launch_chi2(ctx, st, S, Nl, Nb, nblk, blk_off_h, blk_n_h, nmax, th_d, zero_d, data_d, w_d, lb_d, lndet_d,
                                    q_d, chi2_d);
  if (e == cudaSuccess) e = cudaMemcpyAsync(chi2_h, chi2_d, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) {
    cfp_set_error("cfp_cells_chi2: %s", cudaGetErrorString(e));
    return CFP_ERR_CUDA;
  }
  return CFP_OK;
}

extern "C" int cfp_photo_fisher(cfp_ctx* ctx, int Nl, int Nb, int npar, const double* cov_h, const double* dC_h,
                                const double* ell_h, double* fisher_h) {
  CFP_REQUIRE(ctx && cov_h && dC_h && ell_h && fisher_h, "NULL argument");
  CFP_REQUIRE(Nl >= 2 && Nb >= 1 && Nb <= 64 && npar >= 1, "bad dimensions (at most 64 tomographic rows)");
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nbins = Nl - 1;
  double *cov_d = nullptr, *dC_d = nullptr, *ell_d = nullptr, *Z_d = nullptr, *Fl_d = nullptr, *F_d = nullptr;
  int rc = cfp_upload<double>(ctx, cov_h, (size_t)Nl * Nb * Nb, &cov_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, dC_h, (size_t)npar * Nl * Nb * Nb, &dC_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, ell_h, (size_t)Nl, &ell_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, nullptr, (size_t)nbins * npar * Nb * Nb, &Z_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, nullptr, (size_t)nbins * npar * npar, &Fl_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, nullptr, (size_t)npar * npar, &F_d);
  if (rc == CFP_OK) {
    const size_t smem = (size_t)Nb * 2 * Nb * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(photo_fisher_user_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      photo_fisher_user_kernel<<<nbins, 256, smem, st>>>(Nl, Nb, npar, cov_d, dC_d, ell_d, Z_d, Fl_d);
      photo_reduce_kernel<<<(npar * npar + 7) / 8, 256, 0, st>>>(Fl_d, nbins, npar * npar, F_d);
      ctx->launches += 2;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(fisher_h, F_d, (size_t)npar * npar * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      cfp_set_error("cfp_photo_fisher: %s", cudaGetErrorString(e));
      rc = CFP_ERR_CUDA;
    }
  }
  for (void* q : {(void*)cov_d, (void*)dC_d, (void*)ell_d, (void*)Z_d, (void*)Fl_d, (void*)F_d}) cfp_dev_free(ctx, q);
  return rc;
}

extern "C" int cfp_photo_plan_fetch_windows(cfp_photo_plan* pl, double* U_h, double* V_h) {
  CFP_REQUIRE(pl && U_h && V_h, "NULL argument");
  if (!pl->ran) {
    cfp_set_error("cfp_photo_plan_fetch_windows: plan has not been run");
    return CFP_ERR_STATE;
  }
  CFP_CUDA(cudaSetDevice(pl->ctx->device));
  CFP_CUDA(cudaEventSynchronize(pl->run1));
  const int nrow_pad = pl->nb * 8;
  std::vector<double> tmp((size_t)pl->S * nrow_pad * pl->ldz);
  for (int which = 0; which < 2; ++which) {
    CFP_CUDA(cudaMemcpy(tmp.data(), which == 0 ? pl->U_d : pl->V_d, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    double* dst = which == 0 ? U_h : V_h;
    for (int s = 0; s < pl->S; ++s)
      for (int a = 0; a < pl->Nb; ++a)
        for (int z = 0; z < pl->Nz; ++z)
          dst[((size_t)s * pl->Nb + a) * pl->Nz + z] = tmp[((size_t)s * nrow_pad + a) * pl->ldz + z];
  }
  return CFP_OK;
}

