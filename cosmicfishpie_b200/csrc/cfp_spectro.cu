// Spectroscopic galaxy clustering: P_obs(k,mu,z) lattice, derivatives and Fisher reduction.
//
// Kernels (all FP64):
//   spectro_prepare_kernel   blockIdx.y < 3, per (cosmology, z-bin, table): contracts the bicubic B-spline of
//                            P_lin(z,k) / f(z,k) with the z-basis at the bin redshift and converts the resulting
//                            1-D cubic B-spline to piecewise-polynomial (Taylor) form per knot interval; converts
//                            the no-wiggle samples to (value, slope) pairs.  Output lives in HBM, is a few MB and
//                            is read by the main kernel through L1/L2.
//                            blockIdx.y == 3, per z-bin (spectro_setup_block): derived scalars, table pointers,
//                            stencil feed of every (stencil point, bin) and the bin's work list by class.
//   spectro_fisher_kernel    one CTA = one z-bin x a run of tiles (256 wavenumbers x one mu row) walked down the
//                            mu rows of a k block.  Per tile a few threads stage the row constants of every
//                            stencil point (the AP remap is linear in k along a row); each thread then evaluates
//                            ln P_obs of its cell at every stencil point (Kaiser, FoG, dewiggling, spec-z error;
//                            three table look-ups with hints that survive from tile to tile), accumulates the
//                            derivative stencils into a shared-memory tile d[par][cell], and each warp contracts
//                            its 32 cells into the Gram matrix F_ij += w d_i d_j with FP64 tensor-core MMAs
//                            (mma.sync m8n8k4 = DMMA; tcgen05 has no f64 kind).  Per-CTA partials go to HBM.
//   spectro_reduce_kernel    fixed-tree sum of the CTA partials -> F[Z][npar][npar] (deterministic).
//   spectro_chi2_kernel      wedge likelihood of a batch of points: the same evaluation walking down mu columns
//                            against a device-resident data lattice; spectro_data_kernel builds that lattice
//                            from stencil point 0 when the caller has none.
//
// Reference semantics restated here: cosmicfishpie/LSSsurvey/spectro_obs.py:342-933,
// spectro_cov.py:207-228,481-532, fishermatrix/derivatives.py:93-207, cosmicfish.py:495-542.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <math_constants.h>

#include "cfp_common.cuh"

namespace {

constexpr int TILE = 256;       // cells per CTA tile (= threads per CTA)
constexpr int LDPAD = 4;        // row padding of the derivative tile (bank-conflict free DMMA loads)
constexpr int MAX_NB = 6;       // npar <= 48
constexpr double INV_8PI2 = 0.012665147955292222;  // 1/(8 pi^2)
constexpr int CUBIC_BLK = 6;    // doubles per cubic piece:  [lo, hi, c0, c1, c2, c3]
constexpr int LIN_BLK = 4;      // doubles per linear piece: [lo, hi, y, slope]

// Everything the lattice loop needs for one (stencil point, z-bin), derived once by
// spectro_setup_kernel instead of once per thread per cell.
struct alignas(16) SDev {
  // hot fields first, in 16-byte pairs so that the lattice loop fetches them with 128-bit shared loads
  double bias, baos;                 // galaxy bias; bao * invs8sq
  double xminA, xmaxA;               // knot range of the P_lin table (FITPACK clamps the argument to it)
  const double *blkA, *blkB;         // piece tables of P_lin(k), f(k) at this (cosmology, z)
  const double* blkN;                // piece table of P_nw(k)
  int nA, nN;                        // number of pieces
  double w0;                         // weight of this point in the derivative of parameter par0
  int flags, par0;                   // SD_*; parameter fed with weight w0 (-1: none)
  int csr0, csr1;                    // further (parameter, weight) pairs in the CSR arrays
  int nB, sidx;                      // sidx: index of the point in the job's stencil (shared memory holds work order)
  double xminB, xmaxB;
  // row-constant inputs and the rest
  double rqpar, rqper, dzH, A1, A2, sigmap, I0, I12, svr2, fs8, invs8sq, khfac, bao, pshot;
};
enum {
  SD_OWN = 1,           // evaluated (not an alias of the fiducial in this bin)
  SD_QBIAS = 2,
  SD_SHARED_KNOTS = 4,  // P_lin and f share their k knots: one piece search serves both
  SD_SRC = 8,           // P_obs itself is needed (fiducial, or the point the Ps_* derivative reads)
  SD_PSHOT = 16,
  SD_BIASONLY = 32,     // differs from the fiducial only in the galaxy bias: re-uses the fiducial's table look-ups
  SD_FID = 64,          // stencil point 0
  SD_PSSRC = 128,       // the point whose P_obs feeds d lnP / d Ps_i
  SD_NOCLAMP = 256,     // k' of the whole lattice lies inside the knot range of both tables: no clamping needed
  SD_PLAIN = 512        // eval_x_plain applies (see spectro_setup_block)
};

// What one (stencil point, z-bin) needs per mu ROW of the lattice: with the Alcock-Paczynski remap
// k' = k G(mu), mu' = mu'(mu) every mu-dependent factor of P_obs is constant along a row, so the cell loop
// is left with k' = k G, the three table look-ups, one exp, one division and one log.
struct RowC {
  double G;     // k' / k
  double fmu2;  // fs8 * mu'^2              (Kaiser: bias + f(k') * fmu2)
  double errc;  // err / k                  (spectroscopic redshift error exponent = (k errc)^2)
  double sv2;   // (I0 + mu'^2 I12) svr2    (dewiggling damping exponent = sv2 k'^2)
  double fogc;  // mu' sigma_p              (Lorentzian FoG: 1 + (k' fogc)^2)
  double pad;
};

// Plan of the pair kernel (spectro_fisher_pairs_kernel) for one bin, built by spectro_setup_block: which stencil
// point every lane group evaluates.  q = sign * 8 + slot; slots 0..6: parameters whose two stencil points need a full
// evaluation each (weights +1 / -1); slot 7: the bin's bias-only pair (or the fiducial twice).
struct PairPlan {
  int ok;        // the pair kernel applies to this bin (otherwise spectro_fisher_kernel runs it)
  int ps_sel;    // P_obs the shot-noise derivative divides by: 0 fiducial, 1 / 2 the + / - point of slot 7
  int pt[16];    // stencil point of q
  int par[9];    // parameter of slot 0..7; [8]: the bin's shot-noise parameter; -1: none
  int pad;
  double coef[8];  // 1 / stencil denominator of the slot's parameter (0: unused slot)
  double bias0;    // galaxy bias of the fiducial
};

struct SpectroParams {
  int S, Z, Nmu, Nk, npar, nb, nnw, NC;
  unsigned flags;
  int ps_src;
  int materialize;
  int maxint;
  int64_t cell_begin, cell_end;
  const double *k, *mu, *smu, *wk, *wmu;
  const double* scal;     // [S][Z][NSCAL]
  const int* alias;       // [S][Z]
  const int* st_ptr;      // CSR over stencil points: parameters fed by point s
  const int* st_par;
  const double* st_w;
  const double* st_den;   // [npar]
  const double* st_rden;  // [npar] 1 / st_den
  double kmin, kmax;      // range of the wavenumber grid
  const int* ps_bin;      // [npar]
  const int* dead;        // [npar][Z]
  const double *nbar, *vol;
  const int64_t* desc;    // bank descriptors
  int n_tables, tab_pl, tab_f;
  const double* blob;
  double* blk;            // [(c*Z+zb)*2+tab][maxint][CUBIC_BLK]
  double* blknw;          // [(c*Z+zb)][nnw-1][LIN_BLK]
  int* shared_tmp;          // [Z][NC]: 1 if P_lin and f of the cosmology share their k-knots
  SDev* sdev;             // [Z][S]
  int* work;              // [Z][S] stencil points to visit per bin, in order
  int* nwork;             // [Z][8]: 0 full, 1 bias-only, 2 alias points in the bin's work list (after the fiducial);
                          //         3: how many of the full points, listed first, are plain (the specialised loop);
                          //         5: the fiducial is plain as far as its evaluation goes
  double* partial;        // [Z][gridDim.x][T*64]
  double *lnp, *derivs, *veff;
  PairPlan* pairs;        // [Z]
};

// ------------------------------------------------------------------ prepare
struct PrepParams {
  int NC, Z, nnw, maxint, maxncy;
  const double* blob;
  const int64_t* desc;  // [NC][n_tables][5]
  int n_tables;
  int tab[2];
  const double* zbin;   // [Z]
  double* dcoef;        // scratch [(c*Z+zb)*2+tab][maxncy]
  double* blk;
  const double *pnw_k, *pnw_p;
  double* blknw;
};

struct SpectroParams;
__device__ void spectro_setup_block(const SpectroParams& P, int zb);

// blockIdx.y = 0, 1, 2: piece tables of P_lin, f, P_nw for job blockIdx.x = (cosmology, bin);
// blockIdx.y = 3: the per-bin setup (spectro_setup_block) for bin blockIdx.x -- one launch does both.
__global__ void spectro_prepare_kernel(PrepParams P, const SpectroParams SP, int Zbins) {
  if (blockIdx.y == 3) {
    if ((int)blockIdx.x < Zbins) spectro_setup_block(SP, blockIdx.x);
    return;
  }
  const int job = blockIdx.x;  // (c*Z + zb)
  const int c = job / P.Z, zb = job % P.Z;
  const int tab = blockIdx.y;  // 0: P_lin, 1: f, 2: no-wiggle
  if (tab == 2) {
    const double* xk = P.pnw_k + (size_t)c * P.nnw;
    const double* yp = P.pnw_p + (size_t)job * P.nnw;
    double* o = P.blknw + (size_t)job * (P.nnw - 1) * LIN_BLK;
    for (int l = threadIdx.x; l < P.nnw - 1; l += blockDim.x) {
      const double y0 = yp[l];
      o[LIN_BLK * l + 0] = xk[l];
      o[LIN_BLK * l + 1] = xk[l + 1];
      o[LIN_BLK * l + 2] = y0;
      o[LIN_BLK * l + 3] = (yp[l + 1] - y0) / (xk[l + 1] - xk[l]);
    }
    return;
  }
  const int64_t* d = P.desc + ((size_t)c * P.n_tables + P.tab[tab]) * CFP_BANK_DESC;
  const int nx = (int)d[0], ny = (int)d[1];
  const double* tx = P.blob + d[2];
  const double* ty = P.blob + d[3];
  const double* cf = P.blob + d[4];
  const int ncy = ny - 4;
  double z = P.zbin[zb];
  z = fmin(fmax(z, tx[3]), tx[nx - 4]);
  const int lx = cfp_find_interval(tx, nx, z);
  double hx[4];
  cfp_bspl3(tx, lx, z, hx);
  double* dc = P.dcoef + ((size_t)job * 2 + tab) * P.maxncy;
  for (int j = threadIdx.x; j < ncy; j += blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += hx[i] * cf[(size_t)(lx - 3 + i) * ncy + j];
    dc[j] = s;
  }
  __syncthreads();
  // piece l (global knot index l = 3 + m): polynomial in u = k - ty[l]
  const int nint = ny - 7;
  double* blk = P.blk + ((size_t)job * 2 + tab) * P.maxint * CUBIC_BLK;
  for (int m = threadIdx.x; m < nint; m += blockDim.x) {
    const int l = m + 3;
    const double tl = ty[l];
    // polynomial de Boor: h[i] are cubic polynomials in u (coefficients low -> high)
    double h[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) h[i][q] = 0.0;
    h[0][0] = 1.0;
#pragma unroll
    for (int j = 1; j <= 3; ++j) {
      double hh[3][4];
#pragma unroll
      for (int i = 0; i < j; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) hh[i][q] = h[i][q];
#pragma unroll
      for (int q = 0; q < 4; ++q) h[0][q] = 0.0;
#pragma unroll
      for (int i = 0; i < j; ++i) {
        const double tli = ty[l + i + 1], tlj = ty[l + i + 1 - j];
        const double inv = (tli == tlj) ? 0.0 : 1.0 / (tli - tlj);
        const double a0 = (tli - tl) * inv;  // (tli - x)/(tli-tlj) = a0 - inv*u
        const double b0 = (tl - tlj) * inv;  // (x - tlj)/(tli-tlj) = b0 + inv*u
        double nxt[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double lowq = (q > 0) ? hh[i][q - 1] : 0.0;
          h[i][q] += hh[i][q] * a0 - lowq * inv;
          nxt[q] = hh[i][q] * b0 + lowq * inv;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) h[i + 1][q] = nxt[q];
      }
    }
    double out[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double cj = dc[l - 3 + i];
#pragma unroll
      for (int q = 0; q < 4; ++q) out[q] += cj * h[i][q];
    }
    double* o = blk + (size_t)m * CUBIC_BLK;
    o[0] = tl;
    o[1] = ty[l + 1];
#pragma unroll
    for (int q = 0; q < 4; ++q) o[2 + q] = out[q];
  }
}

// one thread per (z-bin, stencil point): derived scalars, table pointers, stencil feed; thread 0 of
// each bin then builds the bin's work list
__device__ void spectro_setup_block(const SpectroParams& P, int zb) {
  extern __shared__ int cls_sm[];                   // [S] unless in batch mode
  const bool batch = (P.materialize & 16) != 0;     // bit 4: batch (likelihood) mode needs no work list
  // do P_lin(z,k) and f(z,k) share their k-knots? (one piece search serves both)  Answered here from the bank
  // itself, so that this block does not wait for the piece tables the other blocks of the launch are writing.
  int* same = P.shared_tmp + zb * P.NC;
  for (int c = threadIdx.x; c < P.NC; c += blockDim.x) {
    const int64_t* dA = P.desc + ((size_t)c * P.n_tables + P.tab_pl) * CFP_BANK_DESC;
    const int64_t* dB = P.desc + ((size_t)c * P.n_tables + P.tab_f) * CFP_BANK_DESC;
    same[c] = (dA[1] == dB[1]);
  }
  __syncthreads();
  {
    // all (cosmology, knot) comparisons as one flat loop: independent loads, a mismatch clears the flag
    int nk_max = 0;
    for (int c = 0; c < P.NC; ++c)
      nk_max = max(nk_max, (int)P.desc[((size_t)c * P.n_tables + P.tab_pl) * CFP_BANK_DESC + 1]);
    for (int x = threadIdx.x; x < P.NC * nk_max; x += blockDim.x) {
      const int c = x / nk_max, i = x - c * nk_max;
      const int64_t* dA = P.desc + ((size_t)c * P.n_tables + P.tab_pl) * CFP_BANK_DESC;
      const int64_t* dB = P.desc + ((size_t)c * P.n_tables + P.tab_f) * CFP_BANK_DESC;
      if (i < (int)dA[1] && dA[1] == dB[1] && P.blob[dA[3] + i] != P.blob[dB[3] + i]) same[c] = 0;
    }
  }
  __syncthreads();
  const int c_fid = (int)P.scal[(size_t)zb * CFP_SP_NSCAL + CFP_SP_COSMO];
  const int nA_fid = (int)P.desc[((size_t)c_fid * P.n_tables + P.tab_pl) * CFP_BANK_DESC + 1] - 7;
  for (int s = threadIdx.x; s < P.S; s += blockDim.x) {
    const double* sc = P.scal + ((size_t)s * P.Z + zb) * CFP_SP_NSCAL;
    SDev o;
    const int c = (int)sc[CFP_SP_COSMO];
    const double qpar = sc[CFP_SP_QPAR], qper = sc[CFP_SP_QPER];
    o.rqpar = 1.0 / qpar;
    o.rqper = 1.0 / qper;
    o.dzH = sc[CFP_SP_DZ] * (1.0 / sc[CFP_SP_HUBBLE]);
    o.bias = sc[CFP_SP_BIAS];
    o.A1 = sc[CFP_SP_A1];
    o.A2 = sc[CFP_SP_A2];
    o.sigmap = sc[CFP_SP_SIGMAP];
    o.I0 = sc[CFP_SP_I0];
    o.I12 = 2.0 * sc[CFP_SP_I1] + sc[CFP_SP_I2];
    o.svr2 = sc[CFP_SP_SVRESCALE] * sc[CFP_SP_SVRESCALE];
    o.fs8 = sc[CFP_SP_FS8];
    o.invs8sq = sc[CFP_SP_INVS8SQ];
    o.khfac = sc[CFP_SP_KHFAC];
    o.pshot = sc[CFP_SP_PSHOT];
    o.bao = (P.flags & CFP_SPF_AP) ? 1.0 / (qper * qper * qpar) : 1.0;
    const int64_t* dA = P.desc + ((size_t)c * P.n_tables + P.tab_pl) * CFP_BANK_DESC;
    const int64_t* dB = P.desc + ((size_t)c * P.n_tables + P.tab_f) * CFP_BANK_DESC;
    const size_t job = (size_t)c * P.Z + zb;
    o.blkA = P.blk + (job * 2 + 0) * P.maxint * CUBIC_BLK;
    o.blkB = P.blk + (job * 2 + 1) * P.maxint * CUBIC_BLK;
    o.blkN = P.blknw + job * (P.nnw - 1) * LIN_BLK;
    o.nA = (int)dA[1] - 7;
    o.nB = (int)dB[1] - 7;
    o.nN = P.nnw - 1;
    o.xminA = P.blob[dA[3] + 3];  // = lo of the first piece, hi of the last (fpbisp.f clamps to [t[3], t[n-4]])
    o.xmaxA = P.blob[dA[3] + dA[1] - 4];
    o.xminB = P.blob[dB[3] + 3];
    o.xmaxB = P.blob[dB[3] + dB[1] - 4];
    o.baos = o.bao * o.invs8sq;
    const bool shared = P.shared_tmp[zb * P.NC + c] != 0;
    const bool own = (s == 0) || (P.alias[s * P.Z + zb] == s);
    o.flags = (own ? SD_OWN : 0) | ((o.A1 != 0.0 || o.A2 != 0.0) ? SD_QBIAS : 0) | (shared ? SD_SHARED_KNOTS : 0) |
              ((s == 0 || s == P.ps_src) ? SD_SRC : 0) | (o.pshot != 0.0 ? SD_PSHOT : 0) | (s == 0 ? SD_FID : 0) |
              (s == P.ps_src ? SD_PSSRC : 0);
    if (s != 0 && own) {
      // same cosmology and same scalars as the fiducial except (possibly) the bias?
      const double* s0 = P.scal + (size_t)zb * CFP_SP_NSCAL;
      bool same = true;
      for (int q = 0; q < CFP_SP_NSCAL; ++q)
        if (q != CFP_SP_BIAS) same = same && (sc[q] == s0[q]);
      if (same) o.flags |= SD_BIASONLY;
    }
    o.sidx = s;
    {
      // k' = k khfac sqrt(mu^2 rqpar^2 + (1 - mu^2) rqper^2) lies between k khfac min(rq) and k khfac max(rq)
      const double gmin = o.khfac * fmin(o.rqpar, o.rqper) * (1.0 - 1e-12);
      const double gmax = o.khfac * fmax(o.rqpar, o.rqper) * (1.0 + 1e-12);
      const double g_lo = (P.flags & CFP_SPF_AP) ? gmin : o.khfac, g_hi = (P.flags & CFP_SPF_AP) ? gmax : o.khfac;
      const double klo = P.kmin * g_lo, khi = P.kmax * g_hi;
      if (klo > fmax(o.xminA, o.xminB) && khi < fmin(o.xmaxA, o.xmaxB)) o.flags |= SD_NOCLAMP;
    }
    // stencil feed: first live (parameter, weight) inline, the rest through the CSR arrays
    o.par0 = -1;
    o.w0 = 0.0;
    o.csr0 = o.csr1 = 0;
    const int j0 = P.st_ptr[s], j1 = P.st_ptr[s + 1];
    int jj = j0;
    for (; jj < j1; ++jj)
      if (!P.dead[P.st_par[jj] * P.Z + zb]) {
        o.par0 = P.st_par[jj];
        o.w0 = P.st_w[jj];
        ++jj;
        break;
      }
    o.csr0 = jj;
    o.csr1 = j1;
    {
      // a "plain" evaluation is what a forecast or a likelihood batch consists of almost entirely: non-linear spectrum
      // with FoG, tables on shared knots covering the whole lattice, no Q-bias, no shot-noise term -- eval_x_plain
      // carries none of those branches
      const bool plain_eval = !(P.flags & CFP_SPF_LINEAR) && (P.flags & CFP_SPF_FOG) &&
                              (o.flags & (SD_SHARED_KNOTS | SD_NOCLAMP)) == (SD_SHARED_KNOTS | SD_NOCLAMP) &&
                              !(o.flags & (SD_QBIAS | SD_PSHOT)) &&
                              o.nA == nA_fid;  // (a piece hint of the fiducial's table is a valid index of this one)
      if (plain_eval) o.flags |= SD_PLAIN;
    }
    P.sdev[(size_t)zb * P.S + s] = o;
    if (!batch) {  // class of the point in the bin's work list (-1: not visited)
      const bool needed = (s == P.ps_src) || (P.materialize & CFP_MAT_LNP) || (o.par0 >= 0);
      const bool plain_eval = (o.flags & SD_PLAIN) != 0;
      // a plain POINT of the Fisher work list also feeds exactly one parameter and nobody reads its P_obs.  (A third loop for the plain point whose P_obs
      // the shot-noise derivatives read was measured slower: 0.813 vs 0.789 ms, register allocation.)
      const bool plain = plain_eval && o.par0 >= 0 && o.csr0 == o.csr1 && !(o.flags & (SD_SRC | SD_PSSRC));
      cls_sm[s] = !needed ? -1 : (!(o.flags & SD_OWN) ? 3 : ((o.flags & SD_BIASONLY) ? 2 : (plain ? 0 : 1)));
      if (s == 0) P.nwork[zb * 8 + 5] = plain_eval ? 1 : 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && !batch) {
    // work list of the bin: [fiducial | full evaluations, plain ones first | bias-only | aliases of the fiducial
    // that feed a derivative or are materialised]; the lattice loop runs one branch-free loop per class
    int n = 0, cnt[4] = {0, 0, 0, 0};
    P.work[zb * P.S + n++] = 0;
    for (int cls = 0; cls < 4; ++cls)
      for (int s = 1; s < P.S; ++s)
        if (cls_sm[s] == cls) P.work[zb * P.S + n++] = s, ++cnt[cls];
    P.nwork[zb * 8 + 0] = cnt[0] + cnt[1];
    P.nwork[zb * 8 + 1] = cnt[2];
    P.nwork[zb * 8 + 2] = cnt[3];
    P.nwork[zb * 8 + 3] = cnt[0];
    // Can the pair kernel take this bin?  A 3PT forecast of plain points: at most 7 parameters with a fully evaluated
    // +/- pair, one bias-only +/- pair, nothing else but the point the shot-noise derivatives read.
    const SDev* sd = P.sdev + (size_t)zb * P.S;
    PairPlan pp;
    bool ok = P.materialize == 0 && P.nwork[zb * 8 + 5] != 0 && P.npar <= MAX_NB * 8;
    pp.ps_sel = 0;
    pp.pad = 0;
    pp.bias0 = sd[0].bias;
    int have[16];
    for (int q = 0; q < 16; ++q) pp.pt[q] = 0, have[q] = 0;
    for (int a = 0; a < 9; ++a) pp.par[a] = -1;
    for (int a = 0; a < 8; ++a) pp.coef[a] = 0.0;
    int nslot = 0;
    for (int s = 1; s < P.S && ok; ++s) {
      const int c = cls_sm[s];
      if (c < 0) continue;
      const SDev& o = sd[s];
      if (c == 3) {  // an alias of the fiducial: only as the point the shot-noise derivatives read
        if (!(s == P.ps_src && o.par0 < 0)) ok = false;
        continue;
      }
      if (c == 1 || !(o.flags & SD_PLAIN) || o.par0 < 0 || o.csr0 != o.csr1 || (o.w0 != 1.0 && o.w0 != -1.0)) {
        ok = false;
        break;
      }
      int slot = -1;
      if (c == 2) {
        slot = 7;
        if (pp.par[7] >= 0 && pp.par[7] != o.par0) ok = false;
      } else {
        for (int a = 0; a < nslot; ++a)
          if (pp.par[a] == o.par0) slot = a;
        if (slot < 0) {
          if (nslot == 7) {
            ok = false;
            break;
          }
          slot = nslot++;
        }
      }
      const int q = (o.w0 == 1.0 ? 0 : 8) + slot;
      if (have[q]) ok = false;
      have[q] = 1;
      pp.pt[q] = s;
      pp.par[slot] = o.par0;
      pp.coef[slot] = P.st_rden[o.par0];
      if (s == P.ps_src) {
        if (c == 2)
          pp.ps_sel = (o.w0 == 1.0) ? 1 : 2;
        else
          ok = false;
      }
    }
    for (int a = 0; a < 8; ++a)
      if (have[a] != have[8 + a]) ok = false;
    for (int par = 0; par < P.npar; ++par)
      if (P.ps_bin[par] == zb) {
        if (pp.par[8] >= 0) ok = false;
        pp.par[8] = par;
      }
    pp.ok = ok ? 1 : 0;
    P.pairs[zb] = pp;
  } else if (threadIdx.x == 0) {
    P.pairs[zb].ok = 0;
  }
}


// ------------------------------------------------------------------ branch-free FP64 math
// The lattice loop only ever needs exp of a non-positive bounded argument, log of a positive normal
// number and 1/x, 1/sqrt(x) of positive normal numbers.  libdevice's general-purpose versions carry
// special-case branches and integer fix-up code that cost more issue slots than the FP64 work itself
// (profiles/r1b: FP64 = 27% of issued instructions).  These versions are accurate to ~1 ulp on the
// stated domains (tested against the CPU oracle at 1e-12 on Fisher elements).
__device__ __forceinline__ double fast_rcp(double x) {  // x > 0, normal
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

__device__ __forceinline__ double fast_div(double a, double b) {  // b > 0, normal
  const double r = fast_rcp(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

__device__ __forceinline__ double fast_rsqrt(double x) {  // x > 0, normal
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);  // 0.5 - 0.5 x y^2
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  return fma(y, e, y);
}

// Constants live in constant memory: a 64-bit literal costs two UMOV issue slots per use on sm_100a
// (profiles/r1c: 45 UMOV per cell evaluation), a constant-bank operand costs one uniform load at most.
// Table-driven versions for the lattice loops: shorter dependency chains and fewer constants than the pure
// polynomial forms above (exp: 32-entry 2^(j/32) table + degree-6 polynomial; log: 64-entry (1/c, -ln(1/c)) table +
// degree-7 log1p).  The tables live in shared memory, filled once per CTA by math_tables_fill.
struct MathTab {
  double e2[32];   // 2^(j/32)
  double2 lg[64];  // {1/c_i, ln c_i},  c_i = 1 + (i + 1/2)/64
};

__device__ __forceinline__ void math_tables_fill(MathTab& t, int tid, int nthreads) {
  for (int j = tid; j < 32; j += nthreads) t.e2[j] = exp2((double)j * (1.0 / 32.0));
  for (int i = tid; i < 64; i += nthreads) {
    const double rc = 1.0 / (1.0 + ((double)i + 0.5) * (1.0 / 64.0));
    t.lg[i] = make_double2(rc, -log(rc));
  }
}

__constant__ double CFP_TEXP[9] = {46.166241308446828384,        // 32 / ln 2
                                   6755399441055744.0,           // 1.5 * 2^52
                                   -2.16608493865351192653e-02,  // -(ln 2 / 32) high part
                                   -5.96317165397058656257e-12,  // -(ln 2 / 32) low part
                                   1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, -700.0};
__constant__ double CFP_TLOG[8] = {1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0,
                                   6.93147180369123816490e-01, 1.90821492927058770002e-10, 0.0};

__device__ __forceinline__ double tab_exp_neg(const MathTab& mt, double x) {  // x <= 0
  x = x < CFP_TEXP[8] ? CFP_TEXP[8] : x;
  double t = fma(x, CFP_TEXP[0], CFP_TEXP[1]);  // low word = round(32 x / ln 2) = 32 n + j
  const int m = __double2loint(t);
  t -= CFP_TEXP[1];
  double r = fma(t, CFP_TEXP[2], x);
  r = fma(t, CFP_TEXP[3], r);                   // |r| <= ln2/64
  const double tj = mt.e2[m & 31];
  double p = fma(CFP_TEXP[4], r, CFP_TEXP[5]);
  p = fma(p, r, CFP_TEXP[6]);
  p = fma(p, r, CFP_TEXP[7]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= tj;
  return __hiloint2double(__double2hiint(p) + ((m >> 5) << 20), __double2loint(p));
}

__device__ __forceinline__ double tab_log(const MathTab& mt, double x) {  // x > 0, normal
  const int hi = __double2hiint(x);
  const int e = (hi >> 20) - 1023;
  const double2 cl = mt.lg[(hi >> 14) & 63];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));  // [1, 2)
  const double t = fma(m, cl.x, -1.0);                                                   // |t| < 1/128
  double p = fma(CFP_TLOG[0], t, CFP_TLOG[1]);
  p = fma(p, t, CFP_TLOG[2]);
  p = fma(p, t, CFP_TLOG[3]);
  p = fma(p, t, CFP_TLOG[4]);
  p = fma(p, t, -0.5);
  const double r = fma(p, t * t, t);  // log1p(t)
  const double de = (double)e;
  return fma(de, CFP_TLOG[5], cl.y + fma(de, CFP_TLOG[6], r));
}

// ------------------------------------------------------------------ lattice evaluation
// piece index j in [0, n-1] whose [lo, hi) contains x (clamped at both ends); hint < 0 = unknown.
// Neighbouring stencil points / cells move by at most a piece or two.
template <int STRIDE>
__device__ __forceinline__ int locate(const double* __restrict__ blk, int n, double x, int hint, double& lo) {
  int j;
  double2 b;
  if (hint < 0) {
    int a = 0, c = n - 1;
    while (a < c) {
      const int mid = (a + c + 1) >> 1;
      if (__ldg(blk + (size_t)mid * STRIDE) <= x)
        a = mid;
      else
        c = mid - 1;
    }
    j = a;
    lo = __ldg(blk + (size_t)j * STRIDE);
    return j;
  }
  j = min(hint, n - 1);  // (a hint carried over from a table with more pieces)
  b = __ldg(reinterpret_cast<const double2*>(blk + (size_t)j * STRIDE));
  while (x >= b.y && j < n - 1) {
    ++j;
    b = __ldg(reinterpret_cast<const double2*>(blk + (size_t)j * STRIDE));
  }
  while (x < b.x && j > 0) {
    --j;
    b = __ldg(reinterpret_cast<const double2*>(blk + (size_t)j * STRIDE));
  }
  lo = b.x;
  return j;
}

// The same with a hint that is known to be a valid index and almost always right (the walk down a mu column moves
// k' by a fraction of a piece per step): one load, two compares, the search only off the common path.
template <int STRIDE>
__device__ __forceinline__ int locate_near(const double* __restrict__ blk, int n, double x, int hint, double& lo) {
  const double2 b = __ldg(reinterpret_cast<const double2*>(blk + (size_t)hint * STRIDE));
  if (x >= b.x && x < b.y) {
    lo = b.x;
    return hint;
  }
  return locate<STRIDE>(blk, n, x, hint, lo);
}

__device__ __forceinline__ double cubic_at(const double* __restrict__ blk, int j, double u) {
  const double2 c01 = __ldg(reinterpret_cast<const double2*>(blk + (size_t)j * CUBIC_BLK + 2));
  const double2 c23 = __ldg(reinterpret_cast<const double2*>(blk + (size_t)j * CUBIC_BLK + 4));
  return c01.x + u * (c01.y + u * (c23.x + u * c23.y));
}

// Row constants of one (stencil point, bin) at one mu (spectro_obs.py:441-506: AP remap and redshift error)
__device__ __forceinline__ RowC row_consts(const SDev& p, unsigned flags, double mu, double smu) {
  RowC r;
  const double mpar = mu * p.rqpar;
  double R = 1.0, muap = mu;
  if (flags & CFP_SPF_AP) {
    const double mper = smu * p.rqper;
    const double s2 = mpar * mpar + mper * mper;
    const double ri = fast_rsqrt(s2);
    R = s2 * ri;
    muap = mpar * ri;
  }
  r.G = p.khfac * R;
  const double mu2 = muap * muap;
  r.fmu2 = p.fs8 * mu2;
  r.errc = p.dzH * (((flags & CFP_SPF_KH_BEFORE_ERR) ? p.khfac : 1.0) * mpar);
  r.sv2 = (p.I0 + mu2 * p.I12) * p.svr2;
  r.fogc = muap * p.sigmap;
  r.pad = 0.0;
  return r;
}

// X = P_obs / exp(-err^2) and err^2 (so that ln P_obs = ln X - err^2 needs a single log and no exp).
// X = (bterm + F)^2 W with  bterm = bias * qf (Q-bias factor),  F = f(k') fs8 mu'^2,
// W = bao/s8^2 * dewiggled P(k') / FoG: a point that differs from the fiducial only in its bias re-uses
// qf, F and W of the fiducial (eval_bias_only).
__device__ __forceinline__ double eval_x(const MathTab& mt, const SDev& p, const RowC& rc, unsigned flags, double k,
                                         int& hA, int& hB, int& hN, double& err2, double& qf, double& F, double& W) {
  const double err = k * rc.errc;
  err2 = err * err;
  const double kap = k * rc.G;
  qf = 1.0;
  if (p.flags & SD_QBIAS) qf = fast_div(1.0 + kap * kap * p.A2, 1.0 + kap * p.A1);
  // tabulated P_lin(k'), f(k') (FITPACK clamps the argument to the knot range)
  double lo;
  double x = kap;
  if (!(p.flags & SD_NOCLAMP)) {
    x = x < p.xminA ? p.xminA : x;
    x = x > p.xmaxA ? p.xmaxA : x;
  }
  hA = locate<CUBIC_BLK>(p.blkA, p.nA, x, hA, lo);
  double u = x - lo;
  const double pdd = cubic_at(p.blkA, hA, u);
  if (p.flags & SD_SHARED_KNOTS) {
    hB = hA;
  } else {
    x = kap;
    if (!(p.flags & SD_NOCLAMP)) {
      x = x < p.xminB ? p.xminB : x;
      x = x > p.xmaxB ? p.xmaxB : x;
    }
    hB = locate<CUBIC_BLK>(p.blkB, p.nB, x, hB, lo);
    u = x - lo;
  }
  F = cubic_at(p.blkB, hB, u) * rc.fmu2;  // spectro_obs.py:584-637
  if (!(flags & CFP_SPF_LINEAR)) {
    // dewiggling, spectro_obs.py:699-722, 811-843; degree-1 spline with extrapolation
    hN = locate<LIN_BLK>(p.blkN, p.nN, kap, hN, lo);
    const double2 ys = __ldg(reinterpret_cast<const double2*>(p.blkN + (size_t)hN * LIN_BLK + 2));
    const double pnw = fma(ys.y, kap - lo, ys.x);
    const double e = tab_exp_neg(mt, -rc.sv2 * (kap * kap));
    W = p.baos * fma(e, pdd - pnw, pnw);  // pdd e + pnw (1 - e)
    if (flags & CFP_SPF_FOG) {            // Lorentzian FoG, spectro_obs.py:639-676
      const double t = kap * rc.fogc;
      W = fast_div(W, fma(t, t, 1.0));
    }
  } else {
    W = p.baos * pdd;
  }
  const double kaiser = fma(p.bias, qf, F);
  return (kaiser * kaiser) * W;
}

__device__ __forceinline__ double eval_x(const MathTab& mt, const SDev& p, const RowC& rc, unsigned flags, double k,
                                         int& hA, int& hB, int& hN, double& err2) {
  double qf, F, W;
  return eval_x(mt, p, rc, flags, k, hA, hB, hN, err2, qf, F, W);
}

// eval_x of a "plain" point (see spectro_setup_block): the same operations in the same order, without the
// configuration branches; hA and hN are valid indices (the fiducial of the tile was evaluated first).
__device__ __forceinline__ double eval_x_plain(const MathTab& mt, const SDev& p, const RowC& rc, double k, int& hA,
                                               int& hN, double& err2, double& F, double& W) {
  const double err = k * rc.errc;
  err2 = err * err;
  const double kap = k * rc.G;
  // The hints are almost always right (the walk down a mu column moves k' by a fraction of a piece per step), so
  // every record is fetched AT THE HINT before any bound is checked: one round trip to the tables per evaluation
  // instead of bounds -> coefficients; a wrong hint takes the search and fetches again.
  const double2* ra = reinterpret_cast<const double2*>(p.blkA + (size_t)hA * CUBIC_BLK);
  const double2* rb = reinterpret_cast<const double2*>(p.blkB + (size_t)hA * CUBIC_BLK);
  const double2* rn = reinterpret_cast<const double2*>(p.blkN + (size_t)hN * LIN_BLK);
  double2 bA = __ldg(ra), a01 = __ldg(ra + 1), a23 = __ldg(ra + 2);
  double2 f01 = __ldg(rb + 1), f23 = __ldg(rb + 2);
  double2 bN = __ldg(rn), ys = __ldg(rn + 1);
  if (!(kap >= bA.x && kap < bA.y)) {
    hA = locate<CUBIC_BLK>(p.blkA, p.nA, kap, hA, bA.x);
    ra = reinterpret_cast<const double2*>(p.blkA + (size_t)hA * CUBIC_BLK);
    rb = reinterpret_cast<const double2*>(p.blkB + (size_t)hA * CUBIC_BLK);
    a01 = __ldg(ra + 1), a23 = __ldg(ra + 2);
    f01 = __ldg(rb + 1), f23 = __ldg(rb + 2);
  }
  if (!(kap >= bN.x && kap < bN.y)) {
    hN = locate<LIN_BLK>(p.blkN, p.nN, kap, hN, bN.x);
    ys = __ldg(reinterpret_cast<const double2*>(p.blkN + (size_t)hN * LIN_BLK + 2));
  }
  const double u = kap - bA.x;
  const double pdd = a01.x + u * (a01.y + u * (a23.x + u * a23.y));
  F = (f01.x + u * (f01.y + u * (f23.x + u * f23.y))) * rc.fmu2;
  const double pnw = fma(ys.y, kap - bN.x, ys.x);
  const double e = tab_exp_neg(mt, -rc.sv2 * (kap * kap));
  const double t = kap * rc.fogc;
  W = fast_div(p.baos * fma(e, pdd - pnw, pnw), fma(t, t, 1.0));
  const double kaiser = p.bias + F;  // = fma(bias, 1, F)
  return (kaiser * kaiser) * W;
}

__device__ __forceinline__ double eval_x_plain(const MathTab& mt, const SDev& p, const RowC& rc, double k, int& hA,
                                               int& hN, double& err2) {
  double F, W;
  return eval_x_plain(mt, p, rc, k, hA, hN, err2, F, W);
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Resident CTAs per SM the kernel is compiled for.  The loop is latency-bound (table look-ups, Horner chains).
// Measured on B200 for the 2-block Gram tile before the plain-point loop: 2 CTAs/SM 1.83 ms, 3 CTAs 1.13 ms, 4 CTAs
// (64 registers) 0.97 ms, 5 CTAs (48 registers, spills around the per-tile code) 0.88 ms; with the plain-point loop
// 4 and 5 CTAs tie (0.782 / 0.785 ms) and 3 CTAs lose (0.859).  Evaluating two stencil points per iteration (all
// records requested up front, one repair branch, one basic block of arithmetic) was slower at every occupancy:
// 1.34 ms (5 CTAs, spills), 1.06 (4), 1.00 (3), 1.16 (2 CTAs, 128 registers, no spills).  Shared memory (44 KB per
// CTA) stops at 5.
constexpr int fisher_min_blocks(int nb) { return nb <= 2 ? 5 : (nb == 3 ? 3 : 2); }
constexpr int FISHER_ROWS = 2;  // mu rows per pair of barriers (1: 0.785 ms, 2: 0.776, 3: 0.772 and 1.1 KB more shared memory)
// The lattice of one bin is cut into tiles of TILE consecutive wavenumbers x ONE mu row; a CTA takes a contiguous
// run of tiles in (k block, mu row) order, i.e. it walks down the rows of a k block.  Along that walk k' = k G(mu)
// moves by a fraction of a table piece per step, so every thread keeps valid piece hints from tile to tile, and
// the row constants of all stencil points are staged in shared memory once per tile by a few threads.
template <int NB, int MINB = fisher_min_blocks(NB)>
__global__ void __launch_bounds__(TILE, MINB) spectro_fisher_kernel(SpectroParams P, int row_lo, int nrows) {
  constexpr int T = NB * (NB + 1) / 2;
  constexpr int LDC = TILE + LDPAD;
  extern __shared__ __align__(16) double smem[];
  double* sm_d = smem;                                 // [NB*8][LDC]
  double* sm_w = smem + NB * 8 * LDC;                  // [TILE]
  MathTab& mt = *reinterpret_cast<MathTab*>(sm_w + TILE);  // (at a compile-time offset: no address arithmetic in the loops)
  SDev* sm_s = reinterpret_cast<SDev*>(&mt + 1);       // [S], in work-list order
  RowC* const sm_r0 = reinterpret_cast<RowC*>(sm_s + P.S);  // [ROWS][S], likewise
  if (P.pairs[blockIdx.y].ok) return;  // (uniform) spectro_fisher_pairs_kernel takes this bin
  math_tables_fill(mt, threadIdx.x, TILE);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int zb = blockIdx.y;
  const int nkb = (P.Nk + TILE - 1) / TILE;
  const int64_t ntile = (int64_t)nkb * nrows;
  // equal runs of tiles per CTA.  (Weighting the nearly empty tiles of a ragged last block of wavenumbers -- Nk = 8193
  // leaves one column -- by their busy warps was measured slower: such a tile is no shorter in latency.)
  const int64_t t_begin = ntile * blockIdx.x / gridDim.x, t_end = ntile * (blockIdx.x + 1) / gridDim.x;
  const int64_t lattice = (int64_t)P.Nmu * P.Nk;
  const int n_full = P.nwork[zb * 8 + 0], n_bias = P.nwork[zb * 8 + 1], n_alias = P.nwork[zb * 8 + 2];
  // lnP lattices are written by the general loop only
  const bool special = !(P.materialize & CFP_MAT_LNP);
  const int n_plain = special ? P.nwork[zb * 8 + 3] : 0;
  const bool fid_plain = P.nwork[zb * 8 + 5] != 0;
  {
    // the points of the bin, staged in the ORDER OF THE WORK LIST: the loops below walk sm_s / sm_r linearly
    constexpr int W = (int)(sizeof(SDev) / sizeof(double));
    const int* work = P.work + zb * P.S;
    const int nwords = (1 + n_full + n_bias + n_alias) * W;
    const double* src = reinterpret_cast<const double*>(P.sdev + (size_t)zb * P.S);
    double* dst = reinterpret_cast<double*>(sm_s);
    for (int i = tid; i < nwords; i += TILE) {
      const int t = i / W;
      dst[i] = src[(size_t)work[t] * W + (i - t * W)];
    }
  }
  __syncthreads();

  double acc[T][2];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t][0] = acc[t][1] = 0.0;

  const double nbar = P.nbar[zb], vol2 = 2.0 * P.vol[zb];
  int kb_cur = -1, ki = 0;
  bool col_ok = false;
  double k = 0.0, wcol = 0.0;
  int hA = -1, hB = -1, hN = -1;

  constexpr int ROWS = FISHER_ROWS;  // mu rows (tiles of one block of wavenumbers) per pair of barriers
  for (int64_t tile0 = t_begin; tile0 < t_end;) {
    const int kb = (int)(tile0 / nrows);
    const int row0 = (int)(tile0 - (int64_t)kb * nrows);
    const int nr = (int)min((int64_t)min(ROWS, nrows - row0), t_end - tile0);
    tile0 += nr;
    if (kb != kb_cur) {  // new block of wavenumbers: new k, Simpson weight; the hints restart
      kb_cur = kb;
      ki = kb * TILE + tid;
      col_ok = ki < P.Nk;
      ki = col_ok ? ki : P.Nk - 1;
      k = __ldg(P.k + ki);
      wcol = (k * k) * __ldg(P.wk + ki);
      hA = hB = hN = -1;
    }
    __syncthreads();  // every warp is done with the previous rows' constants
    for (int idx = tid; idx < nr * (n_full + 1); idx += TILE) {  // the fiducial and the fully evaluated points
      const int rr = idx / (n_full + 1), t = idx - rr * (n_full + 1);
      sm_r0[rr * P.S + t] = row_consts(sm_s[t], P.flags, __ldg(P.mu + row_lo + row0 + rr), __ldg(P.smu + row_lo + row0 + rr));
    }
    __syncthreads();
    for (int rr = 0; rr < nr; ++rr) {
    const int row = row_lo + row0 + rr;
    const RowC* const sm_r = sm_r0 + rr * P.S;
    const int64_t cell = (int64_t)row * P.Nk + ki;
    const bool valid = col_ok && cell >= P.cell_begin && cell < P.cell_end;
    if (__any_sync(0xffffffffu, valid)) {  // (uniform per warp; the barriers are per tile)
#pragma unroll 4
    for (int p = 0; p < NB * 8; ++p) sm_d[p * LDC + tid] = 0.0;
    double Psrc = 1.0;
    // ln P_obs (and P_obs where it is needed) from X and err^2, then the point's contribution to the derivatives
    auto finish = [&](const SDev& sp, double X, double err2) {
      const int fl = sp.flags;
      double lnP, Pobs = 0.0;
      if (fl & SD_PSHOT) {
        Pobs = X * tab_exp_neg(mt, -err2) + sp.pshot;
        lnP = Pobs > 0.0 ? tab_log(mt, Pobs) : CUDART_NAN;
      } else {
        // a spectrum that is not positive (tables ending inside the lattice: the degree-1 no-wiggle spline
        // extrapolates below zero) has no logarithm: NaN, as numpy gives the reference (spectro_obs.py:913-933)
        lnP = X > 0.0 ? tab_log(mt, X) - err2 : CUDART_NAN;
        if (fl & SD_SRC) Pobs = X * tab_exp_neg(mt, -err2);
      }
      if (fl & SD_PSSRC) Psrc = Pobs;
      if ((P.materialize & CFP_MAT_LNP) && valid) P.lnp[((size_t)sp.sidx * P.Z + zb) * lattice + cell] = lnP;
      const int par0 = sp.par0;
      if (par0 >= 0) {
        sm_d[par0 * LDC + tid] += sp.w0 * lnP;
        for (int j = sp.csr0; j < sp.csr1; ++j) {
          const int par = __ldg(P.st_par + j);
          if (!__ldg(P.dead + par * P.Z + zb)) sm_d[par * LDC + tid] += __ldg(P.st_w + j) * lnP;
        }
      }
      return Pobs;
    };
    // the fiducial
    double qf0, F0, W0, err20;
    double X0;
    if (fid_plain && hA >= 0 && hN >= 0) {  // (the first tile of a block of wavenumbers has no hints yet)
      qf0 = 1.0;
      X0 = eval_x_plain(mt, sm_s[0], sm_r[0], k, hA, hN, err20, F0, W0);
      hB = hA;
    } else {
      X0 = eval_x(mt, sm_s[0], sm_r[0], P.flags, k, hA, hB, hN, err20, qf0, F0, W0);
    }
    const double P0 = finish(sm_s[0], X0, err20);
    const SDev* sp = sm_s + 1;
    const RowC* rp = sm_r + 1;
    // points with their own cosmology / AP factors: full evaluation; the plain ones (all of them in a forecast
    // but the one the shot-noise derivatives read) first, in the loop without configuration branches
    for (int i = n_plain; i > 0; --i, ++sp, ++rp) {
      double err2;
      const double X = eval_x_plain(mt, *sp, *rp, k, hA, hN, err2);
      const double lnP = X > 0.0 ? tab_log(mt, X) - err2 : CUDART_NAN;
      sm_d[sp->par0 * LDC + tid] += sp->w0 * lnP;
    }
    if (n_plain > 0) hB = hA;  // (shared knots)
    for (int i = n_full - n_plain; i > 0; --i, ++sp, ++rp) {
      double err2;
      const double X = eval_x(mt, *sp, *rp, P.flags, k, hA, hB, hN, err2);
      finish(*sp, X, err2);
    }
    // points that differ from the fiducial only in the bias: the table look-ups are the fiducial's
    for (int i = n_bias; i > 0; --i, ++sp) {
      const double kaiser = fma(sp->bias, qf0, F0);
      finish(*sp, (kaiser * kaiser) * W0, err20);
    }
    // aliases of the fiducial in this bin
    for (int i = n_alias; i > 0; --i, ++sp) finish(*sp, X0, err20);
    // finish the derivatives of the cell
    for (int par = 0; par < P.npar; ++par) {
      const int pb = __ldg(P.ps_bin + par);
      double d;
      if (pb >= 0)
        d = (pb == zb) ? fast_rcp(Psrc) : 0.0;  // spectro_cov.py:510-532
      else
        d = sm_d[par * LDC + tid] * __ldg(P.st_rden + par);  // dead parameters accumulate nothing: exactly 0
      sm_d[par * LDC + tid] = d;
      if ((P.materialize & CFP_MAT_DERIVS) && valid) P.derivs[((size_t)par * P.Z + zb) * lattice + cell] = d;
    }
    {
      // effective volume on the fiducial spectrum and the quadrature weight of the cell
      const double npobs = nbar * P0;
      const double r = fast_div(npobs, 1.0 + npobs);
      const double veff = INV_8PI2 * r * r;  // spectro_cov.py:223-225
      if ((P.materialize & CFP_MAT_DERIVS) && valid) P.veff[(size_t)zb * lattice + cell] = veff;
      // cosmicfish.py:541: k^2 * 2 * V_survey * veff, times the Simpson weights of :510-511
      sm_w[tid] = valid ? (vol2 * veff) * (__ldg(P.wmu + row) * wcol) : 0.0;
    }
    __syncwarp();
    // Gram update on the FP64 tensor cores: this warp's 32 cells, k-steps of 4 cells
#pragma unroll 2
    for (int ks = 0; ks < 8; ++ks) {
      const int c = warp * 32 + ks * 4 + (lane & 3);
      const double w = sm_w[c];
      double a[NB];
#pragma unroll
      for (int ib = 0; ib < NB; ++ib) a[ib] = sm_d[(ib * 8 + (lane >> 2)) * LDC + c];
      int t = 0;
#pragma unroll
      for (int ib = 0; ib < NB; ++ib)
#pragma unroll
        for (int jb = 0; jb <= ib; ++jb) {
          dmma884(acc[t][0], acc[t][1], a[ib], w * a[jb]);
          ++t;
        }
    }
    __syncwarp();
    }  // any valid cell in this warp's part of the tile
    }  // rows of the group
  }
  // cross-warp reduction in fixed order, then one partial per CTA
  __syncthreads();
  double* red = smem;  // [8 warps][T][64]
#pragma unroll
  for (int t = 0; t < T; ++t) {
    red[((size_t)warp * T + t) * 64 + lane * 2 + 0] = acc[t][0];
    red[((size_t)warp * T + t) * 64 + lane * 2 + 1] = acc[t][1];
  }
  __syncthreads();
  double* out = P.partial + ((size_t)zb * gridDim.x + blockIdx.x) * (T * 64);
  for (int e = tid; e < T * 64; e += TILE) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < TILE / 32; ++w) s += red[(size_t)w * T * 64 + e];
    out[e] = s;
  }
}

// F[zb] = sum over the CTAs' partial tiles in a FIXED tree: warp w adds CTAs w, w+32, ... in order, then the 32
// warp sums are added in warp order.  grid (Z, T*64/32), 1024 threads: lane = element, warp = CTA subset, so the
// ~150 partials of an element are fetched by 32 independent loads at a time instead of one dependent chain.
__global__ void __launch_bounds__(1024) spectro_reduce_kernel(const double* __restrict__ partial, int ncta, int nb,
                                                              int npar, const PairPlan* __restrict__ pairs,
                                                              double* __restrict__ F) {
  __shared__ double red[32][33];
  const int zb = blockIdx.x;
  if (pairs[zb].ok) return;  // spectro_reduce_pairs_kernel writes this bin
  const int T = nb * (nb + 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.y * 32 + lane;  // < T*64 (a multiple of 32)
  double s = 0.0;
  for (int c = warp; c < ncta; c += 32) s += partial[((size_t)zb * ncta + c) * (T * 64) + e];
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0) {
    s = 0.0;
#pragma unroll
    for (int w = 0; w < 32; ++w) s += red[w][lane];
    // decode (tile, lane, r) -> (i, j)
    const int t = e / 64, l = (e % 64) / 2, r = e % 2;
    int ib = 0, acc = 0;
    while (acc + ib + 1 <= t) {
      acc += ib + 1;
      ++ib;
    }
    const int jb = t - acc;
    const int i = ib * 8 + (l >> 2), j = jb * 8 + (l & 3) * 2 + r;
    if (i < npar && j < npar && j <= i) {
      F[((size_t)zb * npar + i) * npar + j] = s;
      F[((size_t)zb * npar + j) * npar + i] = s;
    }
  }
}


// ---------------------------------------------------------------------------- pair kernel
// The Fisher lattice pass of a plain 3PT forecast (PairPlan.ok), organised around the fragment layout of the FP64
// tensor-core MMA instead of around a shared-memory derivative tile.  Two kernels:
//
// spectro_weights_kernel (pre-pass, thread = cell walking down a mu column): the fiducial spectrum of every cell ->
//   effective volume x quadrature weight w, and the shot-noise derivative 1 / P_obs; {w, dps} pairs to HBM
//   (16 B per cell, read back once by the pair kernel while still L2-resident).
//
// spectro_fisher_pairs_kernel:
//   lane = (slot = lane >> 2, cell = lane & 3).  A warp owns four neighbouring wavenumbers and walks DOWN their mu
//   column.  Lane (slot, cell) evaluates the +/- stencil pair of parameter `slot` at its cell and forms
//   d_slot = (ln P+ - ln P-) / (2h) from ONE division and an atanh series (ln(X+/X-) = 2 atanh((X+ - X-)/(X+ + X-)),
//   |t| ~ 1e-2: five terms are exact to rounding) -- no logarithm, no exponent extraction.  That value IS the lane's
//   element of the A fragment (8 parameters x 4 cells) of mma.m8n8k4.f64, and times the cell weight its element of
//   the B fragment: the Gram update F_ij += w d_i d_j needs no transposition through shared memory and no barrier.
//   Along a mu column k' = k G(mu) drifts by a fraction of a table piece, so the piece records (P_lin and f cubics,
//   no-wiggle line) of both stencil points live in REGISTERS and are re-fetched only when k' leaves the piece:
//   the steady state issues no table load at all.  The row constants of the 16 (sign, slot) points for every mu row
//   of the slice are staged once per CTA in shared memory ([row][q], 48-byte records: conflict-free 128-bit loads),
//   already multiplied out so that every mu-dependent factor is one FMA in k^2 or k'.
// Work: tasks = (segment of rows) x (group of 4 wavenumbers), dealt round-robin to the warps of the grid in a fixed
// order (deterministic partial sums); one CTA per SM; partial[zb][cta][80].
constexpr int PK_MAXROWS = 136;  // mu rows staged at once (16 x 48 B each)
constexpr int WG_ROWS = 64;      // most mu rows a CTA of the weights pre-pass walks

struct RowP {    // row constants of the pair kernel, per (row, q)
  double G;      // k' / k
  double fmu2;   // fs8 mu'^2
  double errc2;  // (err / k)^2:  err^2 = k^2 errc2
  double nsvg;   // -sv2 G^2:     dewiggling exponent = k^2 nsvg
  double fg2;    // (G fogc)^2:   FoG denominator = 1 + k^2 fg2
  double pad;
};

struct PRec {  // piece records of one stencil point at the lane's k': a_i, c0, sl are pre-multiplied by bao / sigma8^2
  double loA, a0, a1, a2, a3, f0, f1, f2, f3;
  double loN, c0, sl;  // no-wiggle line c0 + sl k'
  unsigned wA, wN;     // high words of the piece widths: k' - lo is inside the piece if its high word is smaller
};

__device__ __forceinline__ void pair_fetch_cubic(PRec& r, const SDev& sp, double kap, int& h) {
  double lo;
  h = locate<CUBIC_BLK>(sp.blkA, sp.nA, kap, h, lo);
  const double2* ra = reinterpret_cast<const double2*>(sp.blkA + (size_t)h * CUBIC_BLK);
  const double2* rb = reinterpret_cast<const double2*>(sp.blkB + (size_t)h * CUBIC_BLK);
  const double2 b = __ldg(ra), a01 = __ldg(ra + 1), a23 = __ldg(ra + 2), f01 = __ldg(rb + 1), f23 = __ldg(rb + 2);
  r.loA = b.x;
  r.wA = (unsigned)__double2hiint(b.y - b.x);
  r.a0 = a01.x * sp.baos, r.a1 = a01.y * sp.baos, r.a2 = a23.x * sp.baos, r.a3 = a23.y * sp.baos;
  r.f0 = f01.x, r.f1 = f01.y, r.f2 = f23.x, r.f3 = f23.y;
}

__device__ __forceinline__ void pair_fetch_line(PRec& r, const SDev& sp, double kap, int& h) {
  double lo;
  h = locate<LIN_BLK>(sp.blkN, sp.nN, kap, h, lo);
  const double2* rn = reinterpret_cast<const double2*>(sp.blkN + (size_t)h * LIN_BLK);
  const double2 b = __ldg(rn), ys = __ldg(rn + 1);
  // degree-1 spline with extrapolation: the end pieces extend (practically) to infinity
  const double lo_v = (h == 0) ? -1e300 : b.x, hi_v = (h == sp.nN - 1) ? 1e300 : b.y;
  r.loN = lo_v;
  r.wN = (unsigned)__double2hiint(hi_v - lo_v);
  r.sl = ys.y * sp.baos;
  r.c0 = fma(-ys.y, b.x, ys.x) * sp.baos;  // y + slope (k' - lo) = (y - slope lo) + slope k'
}

// exp(x), x <= 0: tab_exp_neg with the clamp done on the high word in the integer pipe (for x <= 0 the high word
// grows with |x| as an unsigned integer) and the polynomial in Estrin form (dependency chain 4 instead of 6)
__device__ __forceinline__ double pair_exp_neg(const MathTab& mt, double x) {
  x = __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC085E000u), __double2loint(x));  // >= -700 (about)
  double t = fma(x, CFP_TEXP[0], CFP_TEXP[1]);  // low word = round(32 x / ln 2) = 32 n + j
  const int m = __double2loint(t);
  t -= CFP_TEXP[1];
  double r = fma(t, CFP_TEXP[2], x);
  r = fma(t, CFP_TEXP[3], r);                   // |r| <= ln2/64
  const double tj = mt.e2[m & 31];
  const double r2 = r * r;
  const double q0 = r + 1.0;
  const double q1 = fma(r, CFP_TEXP[7], 0.5);
  const double q2 = fma(r, CFP_TEXP[5], CFP_TEXP[6]);
  // (1/720 with a zero low word: the r^6 term is below 2e-15, 20 mantissa bits are plenty, and a constant without
  // low word is an immediate operand instead of a constant-bank load)
  double p = fma(__hiloint2double(0x3F56C16C, 0), r2, q2);
  p = fma(p, r2, q1);
  p = fma(p, r2, q0);
  p *= tj;
  return __hiloint2double(__double2hiint(p) + ((m >> 5) << 20), __double2loint(p));
}

// numerator N = (b + f mu'^2)^2 * dewiggled P * bao/s8^2 and FoG denominator of X = N / den
__device__ __forceinline__ void pair_eval(const MathTab& mt, const PRec& r, const RowP& rc, double k2, double kap,
                                          double u, double bias, double& N, double& den) {
  const double pdd = fma(u, fma(u, fma(u, r.a3, r.a2), r.a1), r.a0);
  const double pf = fma(u, fma(u, fma(u, r.f3, r.f2), r.f1), r.f0);
  const double pnw = fma(r.sl, kap, r.c0);
  const double e = pair_exp_neg(mt, k2 * rc.nsvg);
  const double dw = fma(e, pdd - pnw, pnw);
  den = fma(k2, rc.fg2, 1.0);
  const double kaiser = fma(pf, rc.fmu2, bias);
  N = (kaiser * kaiser) * dw;
}

__constant__ double CFP_ATANH[4] = {1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0};

// {w, dps}[zb][row - row_lo][k]: quadrature weight x 2 V veff of the cell on the fiducial spectrum, and the
// shot-noise derivative 1 / P_obs (0 if the bin has no shot-noise parameter).  grid (k blocks, Z, row segments);
// a thread walks down `seglen` rows of its wavenumber with the fiducial's piece records in registers.
__global__ void __launch_bounds__(TILE) spectro_weights_kernel(SpectroParams P, int row_lo, int nrows, int seglen,
                                                               double2* __restrict__ wd) {
  const int zb = blockIdx.y;
  if (!P.pairs[zb].ok) return;
  __shared__ SDev sp;
  __shared__ RowP rows[WG_ROWS];
  __shared__ double wmu[WG_ROWS];
  __shared__ MathTab mt;
  const int tid = threadIdx.x;
  if (tid < (int)(sizeof(SDev) / sizeof(double)))
    reinterpret_cast<double*>(&sp)[tid] = reinterpret_cast<const double*>(P.sdev + (size_t)zb * P.S)[tid];
  math_tables_fill(mt, tid, TILE);
  __syncthreads();
  const int r0 = blockIdx.z * seglen, nr = min(seglen, nrows - r0);
  if (tid < nr) {
    const int row = row_lo + r0 + tid;
    const RowC rc = row_consts(sp, P.flags, __ldg(P.mu + row), __ldg(P.smu + row));
    RowP rp;
    rp.G = rc.G, rp.fmu2 = rc.fmu2, rp.errc2 = rc.errc * rc.errc, rp.nsvg = -rc.sv2 * (rc.G * rc.G);
    const double gf = rc.G * rc.fogc;
    rp.fg2 = gf * gf, rp.pad = 0.0;
    rows[tid] = rp;
    wmu[tid] = __ldg(P.wmu + row);
  }
  __syncthreads();
  const int ki = blockIdx.x * TILE + tid;
  const int ldk = ((P.Nk + 3) >> 2) << 2;  // rows padded to whole groups of four wavenumbers (zero weight)
  if (ki >= P.Nk) {
    if (ki < ldk)
      for (int r = 0; r < nr; ++r) wd[((size_t)zb * nrows + (r0 + r)) * ldk + ki] = make_double2(0.0, 0.0);
    return;
  }
  const PairPlan& plan = P.pairs[zb];
  const int ps_sel = plan.ps_sel;
  const bool has_ps = plan.par[8] >= 0;
  const double bias0 = sp.bias;
  const double bsrc = ps_sel == 0 ? bias0 : P.sdev[(size_t)zb * P.S + plan.pt[(ps_sel == 1 ? 0 : 8) + 7]].bias;
  const double nbar = P.nbar[zb], vol2 = 2.0 * P.vol[zb];
  const double k = __ldg(P.k + ki);
  const double k2 = k * k;
  const double wcol = k2 * __ldg(P.wk + ki);
  PRec A;
  A.loA = A.loN = 0.0;
  A.wA = A.wN = 0u;
  A.a0 = A.a1 = A.a2 = A.a3 = A.f0 = A.f1 = A.f2 = A.f3 = A.c0 = A.sl = 0.0;
  int hA = -1, hN = -1;
  double2* out = wd + ((size_t)zb * nrows + r0) * ldk + ki;
  int64_t cell = (int64_t)(row_lo + r0) * P.Nk + ki;
  for (int r = 0; r < nr; ++r, out += ldk, cell += P.Nk) {
    const RowP& rc = rows[r];
    const double kap = k * rc.G;
    double u = kap - A.loA;
    const double v = kap - A.loN;
    if (!((unsigned)__double2hiint(u) < A.wA)) pair_fetch_cubic(A, sp, kap, hA), u = kap - A.loA;
    if (!((unsigned)__double2hiint(v) < A.wN)) pair_fetch_line(A, sp, kap, hN);
    const double pdd = fma(u, fma(u, fma(u, A.a3, A.a2), A.a1), A.a0);
    const double F = fma(u, fma(u, fma(u, A.f3, A.f2), A.f1), A.f0) * rc.fmu2;
    const double pnw = fma(A.sl, kap, A.c0);
    const double dw = fma(pair_exp_neg(mt, k2 * rc.nsvg), pdd - pnw, pnw);
    // X e^{-err^2} / FoG denominator
    const double damp = pair_exp_neg(mt, -(k2 * rc.errc2)) * fast_rcp(fma(k2, rc.fg2, 1.0));
    const double k0 = bias0 + F;
    const double P0 = ((k0 * k0) * dw) * damp;
    const double ks = bsrc + F;
    const double Psrc = ((ks * ks) * dw) * damp;  // (= P0 unless the bias pair's point is the source)
    const double npobs = nbar * P0;
    const double rr = npobs * fast_rcp(1.0 + npobs);
    const double veff = INV_8PI2 * rr * rr;  // spectro_cov.py:223-225
    const bool valid = cell >= P.cell_begin && cell < P.cell_end;
    // cosmicfish.py:541: k^2 * 2 * V_survey * veff, times the Simpson weights of :510-511
    const double w = valid ? (vol2 * veff) * (wmu[r] * wcol) : 0.0;
    *out = make_double2(w, has_ps ? fast_rcp(Psrc) : 0.0);  // spectro_cov.py:510-532
  }
}

template <int PK_WARPS>
__global__ void __launch_bounds__(PK_WARPS * 32, 1)
    spectro_fisher_pairs_kernel(SpectroParams P, int row_lo, int nrows, int srows, int nseg,
                                const double2* __restrict__ wd, double* __restrict__ partial) {
  constexpr int PK_THREADS = PK_WARPS * 32;
  const int zb = blockIdx.y;
  if (!P.pairs[zb].ok) return;  // (uniform: the general kernel takes this bin)
  extern __shared__ __align__(16) double smem[];
  RowP* const sm_rows = reinterpret_cast<RowP*>(smem);                   // [srows][16]
  MathTab& mt = *reinterpret_cast<MathTab*>(sm_rows + (size_t)srows * 16);
  SDev* const sm_pt = reinterpret_cast<SDev*>(&mt + 1);                   // [16]
  PairPlan& plan = *reinterpret_cast<PairPlan*>(sm_pt + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = lane >> 2, cell = lane & 3;
  if (tid < (int)(sizeof(PairPlan) / sizeof(int))) reinterpret_cast<int*>(&plan)[tid] = reinterpret_cast<const int*>(P.pairs + zb)[tid];
  math_tables_fill(mt, tid, PK_THREADS);
  __syncthreads();
  {
    constexpr int W = (int)(sizeof(SDev) / sizeof(double));
    const double* src = reinterpret_cast<const double*>(P.sdev + (size_t)zb * P.S);
    double* dst = reinterpret_cast<double*>(sm_pt);
    for (int i = tid; i < 16 * W; i += PK_THREADS) {
      const int q = i / W;
      dst[i] = src[(size_t)plan.pt[q] * W + (i - q * W)];
    }
  }
  __syncthreads();
  const SDev& spA = sm_pt[slot];
  const SDev& spB = sm_pt[8 + slot];
  const double biasA = spA.bias, biasB = spB.bias;
  const double coef = plan.coef[slot];
  const int nquad = (P.Nk + 3) >> 2;
  const int ldk = nquad << 2;
  // Gram accumulators: the 8 x 8 tile of the slots on the tensor cores; the shot-noise row F[8][slot] and F[8][8]
  // as two scalar FMAs per lane (only one row of a second tile would be non-zero)
  double acc0 = 0.0, acc1 = 0.0, acc8 = 0.0, acc88 = 0.0;

  for (int r_base = 0; r_base < nrows; r_base += srows) {
    const int nr = min(srows, nrows - r_base);
    __syncthreads();
    for (int i = tid; i < nr * 16; i += PK_THREADS) {
      const int rr = i >> 4, q = i & 15;
      const int row = row_lo + r_base + rr;
      const RowC rc = row_consts(sm_pt[q], P.flags, __ldg(P.mu + row), __ldg(P.smu + row));
      RowP rp;
      rp.G = rc.G, rp.fmu2 = rc.fmu2, rp.errc2 = rc.errc * rc.errc, rp.nsvg = -rc.sv2 * (rc.G * rc.G);
      const double gf = rc.G * rc.fogc;
      rp.fg2 = gf * gf, rp.pad = 0.0;
      sm_rows[i] = rp;
    }
    __syncthreads();
    const int segs = (nr == nrows) ? nseg : 1;  // (a slice staged in several pieces is not cut further)
    const int seglen = (nr + segs - 1) / segs;
    const int ntask = nquad * segs;
    for (int task = blockIdx.x * PK_WARPS + warp; task < ntask; task += gridDim.x * PK_WARPS) {
      const int seg = task / nquad, quad = task - seg * nquad;
      const int r0 = seg * seglen, r1 = min(nr, r0 + seglen);
      const int kcol = quad * 4 + cell;  // (columns beyond the lattice carry zero weight in wd)
      const double k = __ldg(P.k + min(kcol, P.Nk - 1));
      const double k2 = k * k;
      PRec A, B;
      A.loA = A.loN = B.loA = B.loN = 0.0;
      A.wA = A.wN = B.wA = B.wN = 0u;  // nothing fetched yet
      A.a0 = A.a1 = A.a2 = A.a3 = A.f0 = A.f1 = A.f2 = A.f3 = A.c0 = A.sl = 0.0;
      B.a0 = B.a1 = B.a2 = B.a3 = B.f0 = B.f1 = B.f2 = B.f3 = B.c0 = B.sl = 0.0;
      int hAa = -1, hAn = -1, hBa = -1, hBn = -1;
      const RowP* rcp = sm_rows + (size_t)r0 * 16 + slot;
      const double2* wdp = wd + ((size_t)zb * nrows + (r_base + r0)) * ldk + kcol;
      // The loop is software-pipelined by one row: iteration r evaluates the pair of row r and FINISHES row r - 1
      // (division, series, Gram update), so that the serial tail of one row overlaps the table polynomials and
      // exponentials of the next in one basic block; {w, dps} of a row are requested one iteration before their use.
      // The first iteration finishes a dummy row of zero weight, the last one (r == r1) re-evaluates row r1 - 1.
      double num_p = 1.0, dnm_p = 1.0, ed_p = 0.0;
      double2 wv_p = make_double2(0.0, 0.0);
      for (int r = r0; r <= r1; ++r) {
        const double2 wv = __ldg(wdp);
        const RowP& rcA = rcp[0];
        const RowP& rcB = rcp[8];
        const double kapA = k * rcA.G, kapB = k * rcB.G;
        double uA = kapA - A.loA, uB = kapB - B.loA;
        const double vA = kapA - A.loN, vB = kapB - B.loN;
        // inside the pieces held in registers?  (high words compared as unsigned integers: a negative or too large
        // offset fails; so does, harmlessly, one within 2^-20 of the piece's upper end)
        const bool okA = (unsigned)__double2hiint(uA) < A.wA, okAn = (unsigned)__double2hiint(vA) < A.wN;
        const bool okB = (unsigned)__double2hiint(uB) < B.wA, okBn = (unsigned)__double2hiint(vB) < B.wN;
        if (!(okA && okAn && okB && okBn)) {
          if (!okA) pair_fetch_cubic(A, spA, kapA, hAa), uA = kapA - A.loA;
          if (!okAn) pair_fetch_line(A, spA, kapA, hAn);
          if (!okB) pair_fetch_cubic(B, spB, kapB, hBa), uB = kapB - B.loA;
          if (!okBn) pair_fetch_line(B, spB, kapB, hBn);
        }
        double NA, denA, NB, denB;
        pair_eval(mt, A, rcA, k2, kapA, uA, biasA, NA, denA);
        pair_eval(mt, B, rcB, k2, kapB, uB, biasB, NB, denB);
        const double ed = k2 * (rcA.errc2 - rcB.errc2);  // err+^2 - err-^2
        // ---- finish the previous row: d = (ln P+ - ln P-) / (2h),  ln P = ln(N / den) - err^2
        {
          const double t = (num_p - dnm_p) * fast_rcp(num_p + dnm_p);
          const double t2 = t * t;
          // (1/9 and 1/7 truncated to a zero low word -- immediates; their terms are below 1e-13 and 1e-10 of t)
          double p = fma(__hiloint2double(0x3FBC71C7, 0), t2, __hiloint2double(0x3FC24925, 0));
          p = fma(p, t2, CFP_ATANH[2]);
          p = fma(p, t2, CFP_ATANH[3]);
          double lnr = fma(p * t2, t + t, t + t);
          // the series needs |t| < 0.03 (t^10 / 11 < 1e-16 relative) and both spectra positive (sign bits clear)
          if (!(fabs(t) < 0.03 && (__double2hiint(num_p) | __double2hiint(dnm_p)) >= 0)) {
            // a spectrum that is not positive has no logarithm: NaN, as numpy gives the reference (spectro_obs.py:913-933)
            lnr = (num_p > 0.0 && dnm_p > 0.0) ? tab_log(mt, num_p) - tab_log(mt, dnm_p) : CUDART_NAN;
          }
          const double d = (lnr - ed_p) * coef;
          const double b0 = wv_p.x * d;
          dmma884(acc0, acc1, d, b0);
          acc8 = fma(wv_p.y, b0, acc8);
          acc88 = fma(wv_p.y * wv_p.y, wv_p.x, acc88);
        }
        num_p = NA * denB, dnm_p = NB * denA, ed_p = ed, wv_p = wv;
        if (r + 1 < r1) rcp += 16, wdp += ldk;  // (the extra iteration stays on the last row)
      }
      // (what is still carried here is the duplicate evaluation of the last row: dropped)
    }
  }
  // the shot-noise row: sum over the four cells of a slot
  acc8 += __shfl_xor_sync(0xffffffffu, acc8, 1);
  acc8 += __shfl_xor_sync(0xffffffffu, acc8, 2);
  acc88 += __shfl_xor_sync(0xffffffffu, acc88, 1);
  acc88 += __shfl_xor_sync(0xffffffffu, acc88, 2);
  // cross-warp reduction in fixed order, then one partial per CTA: [tile 0: 64][F[8][slot]: 8][F[8][8]: 1]
  __syncthreads();
  double* red = smem;  // [PK_WARPS][80]
  red[warp * 80 + lane * 2 + 0] = acc0;
  red[warp * 80 + lane * 2 + 1] = acc1;
  if (cell == 0) red[warp * 80 + 64 + slot] = acc8;
  if (lane == 0) red[warp * 80 + 72] = acc88;
  __syncthreads();
  double* out = partial + ((size_t)zb * gridDim.x + blockIdx.x) * 80;
  for (int e = tid; e < 73; e += PK_THREADS) {
    double sacc = 0.0;
#pragma unroll
    for (int w = 0; w < PK_WARPS; ++w) sacc += red[w * 80 + e];
    out[e] = sacc;
  }
}

// ---------------------------------------------------------------------------- math self-test
// y[i] = f(x[i]) for the branch-free FP64 functions of the lattice loops (tests sweep their stated domains against libm)
__global__ void math_selftest_kernel(int which, int n, const double* __restrict__ x, double* __restrict__ y) {
  __shared__ MathTab mt;
  math_tables_fill(mt, threadIdx.x, blockDim.x);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  double r;
  switch (which) {
    case 0: r = tab_exp_neg(mt, v); break;
    case 1: r = tab_log(mt, v); break;
    case 2: r = fast_rcp(v); break;
    case 3: r = fast_rsqrt(v); break;
    case 4: r = fast_div(1.0, v); break;
    case 5: r = pair_exp_neg(mt, v); break;
    default: r = CUDART_NAN; break;
  }
  y[i] = r;
}

// F[zb] of the bins the pair kernel took: internal order (slots 0..7, shot noise) -> parameters, zeros elsewhere.
// grid (Z, 8); one warp per element of the lower triangle, lanes over the CTA partials (fixed order: lane l adds CTAs
// l, l + 32, ..., then a shuffle tree).
__global__ void __launch_bounds__(1024) spectro_reduce_pairs_kernel(const double* __restrict__ partial, int ncta,
                                                                    int npar, const PairPlan* __restrict__ pairs,
                                                                    double* __restrict__ F) {
  const int zb = blockIdx.x;
  const PairPlan& plan = pairs[zb];
  if (!plan.ok) return;
  __shared__ int inv[MAX_NB * 8];
  for (int p = threadIdx.x; p < npar; p += blockDim.x) {
    int a = -1;
    for (int q = 0; q < 9; ++q)
      if (plan.par[q] == p) a = q;
    inv[p] = a;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, nwarp = (blockDim.x >> 5) * gridDim.y;
  for (int e = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); e < npar * npar; e += nwarp) {
    const int i = e / npar, j = e - i * npar;
    if (j > i) continue;
    const int ai = inv[i], aj = inv[j];
    double s = 0.0;
    if (ai >= 0 && aj >= 0) {
      const int hi = max(ai, aj), lo = min(ai, aj);
      // [tile of the 8 slots: fragment order][F[8][slot]][F[8][8]]
      const int el = hi < 8 ? (hi * 4 + (lo >> 1)) * 2 + (lo & 1) : (lo < 8 ? 64 + lo : 72);
      for (int c = lane; c < ncta; c += 32) s += partial[((size_t)zb * ncta + c) * 80 + el];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if (lane == 0) {
      F[((size_t)zb * npar + i) * npar + j] = s;
      F[((size_t)zb * npar + j) * npar + i] = s;
    }
  }
}

// ---------------------------------------------------------------------------- likelihood (wedges)
// Grid-stride walk over the lattice that keeps (mu index, k index) incrementally -- no 64-bit division per
// cell -- and re-derives the row constants only when the thread's mu row changes.
struct CellWalk {
  int64_t cell;
  int mi, ki, dmi, dki;
  __device__ __forceinline__ CellWalk(int64_t first, int64_t stride, int Nk) {
    cell = first;
    mi = (int)(first / Nk);
    ki = (int)(first - (int64_t)mi * Nk);
    dmi = (int)(stride / Nk);
    dki = (int)(stride - (int64_t)dmi * Nk);
  }
  __device__ __forceinline__ void next(int64_t stride, int Nk) {
    cell += stride;
    mi += dmi;
    ki += dki;
    if (ki >= Nk) {
      ki -= Nk;
      ++mi;
    }
  }
};

// data[zb][cell] = P_obs(stencil point 0) + 1/nbar + shot_data[zb]
__global__ void __launch_bounds__(TILE, 3) spectro_data_kernel(SpectroParams P, const double* __restrict__ shot_data,
                                                               double* __restrict__ data) {
  __shared__ SDev sp;
  __shared__ MathTab mt;
  const int zb = blockIdx.y, tid = threadIdx.x;
  if (tid < (int)(sizeof(SDev) / sizeof(double)))
    reinterpret_cast<double*>(&sp)[tid] = reinterpret_cast<const double*>(P.sdev + (size_t)zb * P.S)[tid];
  math_tables_fill(mt, tid, TILE);
  __syncthreads();
  const int64_t lattice = (int64_t)P.Nmu * P.Nk;
  const double extra = 1.0 / P.nbar[zb] + (shot_data ? shot_data[zb] : 0.0);
  int hA = -1, hB = -1, hN = -1, row = -1;
  RowC rc;
  const int64_t stride = (int64_t)gridDim.x * TILE;
  for (CellWalk w((int64_t)blockIdx.x * TILE + tid, stride, P.Nk); w.cell < lattice; w.next(stride, P.Nk)) {
    if (w.mi != row) {
      row = w.mi;
      rc = row_consts(sp, P.flags, __ldg(P.mu + row), __ldg(P.smu + row));
    }
    double err2;
    const double X = eval_x(mt, sp, rc, P.flags, __ldg(P.k + w.ki), hA, hB, hN, err2);
    data[(size_t)zb * lattice + w.cell] = X * tab_exp_neg(mt, -err2) + sp.pshot + extra;
  }
}

// The factors of P_obs of stencil point `s` as separate lattices (the reference's accessors kaiserTerm, FingersOfGod,
// dewiggled_pdd, spec_err_z and observed_Pgg: spectro_obs.py:476-506, 584-676, 811-911), evaluated with the same
// table look-ups and row constants as the fused kernels.  out[t][zb][cell], t = 0 Kaiser term  b(z) q(k') + f(k') mu'^2
// [x sigma8], 1 Fingers-of-God factor, 2 de-wiggled spectrum [/ sigma8^2], 3 redshift-error factor exp(-err^2 / 2),
// 4 P_obs.
__global__ void __launch_bounds__(TILE, 3) spectro_terms_kernel(SpectroParams P, int s, double* __restrict__ out) {
  __shared__ SDev sp;
  __shared__ MathTab mt;
  const int zb = blockIdx.y, tid = threadIdx.x;
  if (tid < (int)(sizeof(SDev) / sizeof(double)))
    reinterpret_cast<double*>(&sp)[tid] = reinterpret_cast<const double*>(P.sdev + (size_t)zb * P.S + s)[tid];
  math_tables_fill(mt, tid, TILE);
  __syncthreads();
  const int64_t lattice = (int64_t)P.Nmu * P.Nk;
  const size_t plane = (size_t)P.Z * lattice;
  int hA = -1, hB = -1, hN = -1, row = -1;
  RowC rc;
  const int64_t stride = (int64_t)gridDim.x * TILE;
  for (CellWalk w((int64_t)blockIdx.x * TILE + tid, stride, P.Nk); w.cell < lattice; w.next(stride, P.Nk)) {
    if (w.mi != row) {
      row = w.mi;
      rc = row_consts(sp, P.flags, __ldg(P.mu + row), __ldg(P.smu + row));
    }
    const double k = __ldg(P.k + w.ki);
    double err2, qf, F, W;
    const double X = eval_x(mt, sp, rc, P.flags, k, hA, hB, hN, err2, qf, F, W);
    const double kap = k * rc.G;
    const double t = kap * rc.fogc;
    const double fogden = ((P.flags & CFP_SPF_FOG) && !(P.flags & CFP_SPF_LINEAR)) ? fma(t, t, 1.0) : 1.0;
    const size_t o = (size_t)zb * lattice + w.cell;
    out[0 * plane + o] = fma(sp.bias, qf, F);
    out[1 * plane + o] = 1.0 / fogden;
    out[2 * plane + o] = (W * fogden) / sp.bao;  // W = bao / s8^2 * dewiggled / FoG denominator
    out[3 * plane + o] = exp(-0.5 * err2);       // P_obs carries its square, exp(-err^2)
    out[4 * plane + o] = X * tab_exp_neg(mt, -err2) + sp.pshot;
  }
}

// Wedge chi^2 of a batch of points.  grid (gx, chunks): a chunk is up to CHI_CH points of ONE bin whose scalar blocks
// differ only in the galaxy bias (a sampler that draws nuisance parameters around cosmologies of the bank produces
// exactly that): the table look-ups, exp and FoG of a cell are evaluated once per chunk -- X = (b q + F)^2 W, see
// eval_x -- and each member point costs a handful of FP64 operations.
// partial[s][zb][cta] = sum over this CTA's cells of w (P_d - P_t)^2 / P_d^2.
// A CTA owns blocks of TILE consecutive k and walks DOWN the mu rows of each block: along a column k' = k G(mu)
// drifts by a fraction of a piece per step, so the piece hints of the three tables stay valid, loads of the data
// lattice stay coalesced, and the row constants of the chunk (staged once per CTA) are broadcast reads.
constexpr int CHI_ROWS = 256;  // mu rows staged at a time (12 KB of RowC + 2 KB of row weights)
constexpr int CHI_CH = 16;     // points per chunk of a group (one accumulator each in registers)
// CH = 1 (points that share their block with nobody: 5 CTAs/SM) or CHI_CH (3 CTAs/SM)
template <int CH, int MINB>
__global__ void __launch_bounds__(TILE, MINB) spectro_chi2_kernel(SpectroParams P, const double* __restrict__ data,
                                                               const double* __restrict__ shot,
                                                               const int* __restrict__ ch_zb,
                                                               const int* __restrict__ ch_off,
                                                               const int* __restrict__ ch_cnt,
                                                               const int* __restrict__ mem_s,
                                                               double* __restrict__ partial) {
  __shared__ SDev sp;
  __shared__ RowC rows[CHI_ROWS];
  __shared__ double wrow[CHI_ROWS];
  __shared__ double red[CH][TILE / 32];
  __shared__ double mb[CH], mx[CH];  // bias and additive noise of the member points
  __shared__ int ms[CH];
  __shared__ MathTab mt;
  const int tid = threadIdx.x;
  const int zb = ch_zb[blockIdx.y], off = ch_off[blockIdx.y], cnt = ch_cnt[blockIdx.y];
  const int s0 = mem_s[off];  // representative: every member shares its block except the bias
  if (tid < (int)(sizeof(SDev) / sizeof(double)))
    reinterpret_cast<double*>(&sp)[tid] = reinterpret_cast<const double*>(P.sdev + (size_t)zb * P.S + s0)[tid];
  math_tables_fill(mt, tid, TILE);
  __syncthreads();
  if (tid < CH) {
    const int sj = mem_s[off + min(tid, cnt - 1)];
    ms[tid] = sj;
    mb[tid] = P.sdev[(size_t)zb * P.S + sj].bias;
    mx[tid] = 1.0 / P.nbar[zb] + shot[(size_t)sj * P.Z + zb] + sp.pshot;
  }
  const int64_t lattice = (int64_t)P.Nmu * P.Nk;
  const double pref = 2.0 * P.vol[zb] * INV_8PI2;  // spectro_like.py:162-178
  const double* __restrict__ dz = data + (size_t)zb * lattice;
  const int nkb = (P.Nk + TILE - 1) / TILE;
  const bool plain = (sp.flags & SD_PLAIN) != 0;
  double acc[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) acc[j] = 0.0;
  for (int r0 = 0; r0 < P.Nmu; r0 += CHI_ROWS) {
    const int nr = min(CHI_ROWS, P.Nmu - r0);
    __syncthreads();
    for (int r = tid; r < nr; r += TILE) {
      rows[r] = row_consts(sp, P.flags, __ldg(P.mu + r0 + r), __ldg(P.smu + r0 + r));
      wrow[r] = __ldg(P.wmu + r0 + r) * pref;
    }
    __syncthreads();
    for (int kb = blockIdx.x; kb < nkb; kb += gridDim.x) {
      const int ki = kb * TILE + tid;
      if (ki >= P.Nk) continue;
      const double k = __ldg(P.k + ki);
      const double wcol = __ldg(P.wk + ki) * (k * k);
      const double* __restrict__ dcol = dz + (size_t)r0 * P.Nk + ki;
      int hA = -1, hB = -1, hN = -1;
      for (int r = 0; r < nr; ++r) {
        const double Pd = __ldg(dcol + (size_t)r * P.Nk);
        double err2, qf, F, W;
        if (plain && hA >= 0) {  // (every row of a column but the first)
          qf = 1.0;
          eval_x_plain(mt, sp, rows[r], k, hA, hN, err2, F, W);
        } else {
          eval_x(mt, sp, rows[r], P.flags, k, hA, hB, hN, err2, qf, F, W);
        }
        const double We = W * tab_exp_neg(mt, -err2);
        const double rPd = fast_rcp(Pd);
        const double w = wrow[r] * wcol;
#pragma unroll
        for (int j = 0; j < CH; ++j)
          if (j < cnt) {
            const double kaiser = fma(mb[j], qf, F);
            const double Pt = fma(kaiser * kaiser, We, mx[j]);
            const double q = (Pd - Pt) * rPd;
            acc[j] = fma(w, q * q, acc[j]);
          }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    double a = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0) red[j][tid >> 5] = a;
  }
  __syncthreads();
  if (tid < cnt) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < TILE / 32; ++w) t += red[tid][w];
    partial[((size_t)ms[tid] * P.Z + zb) * gridDim.x + blockIdx.x] = t;
  }
}

// Legendre-multipole chi^2 of a batch of points, fused (reference: spectro_like.py:48-56, 125-140, 192-233): a thread
// owns one wavenumber, walks down the mu rows accumulating P_l(k) = sum_mu lw_l(mu) P_t(k, mu) for l = 0, 2, 4 of up
// to LEG_CH points that share everything but the bias, then contracts (P_l,data - P_l)^T C^-1 (P_l,data - P_l) with
// the 3x3 inverse covariance of its (k, bin).  The theory lattices never exist in memory.
// partial[s][zb][cta] = this CTA's share of chi2[s].  Requires Nmu <= CHI_ROWS (all rows staged at once).
constexpr int LEG_CH = 4;
__global__ void __launch_bounds__(TILE, 4) spectro_legendre_chi2_kernel(
    SpectroParams P, const double* __restrict__ pl_data /*[Nk][Z][3]*/, const double* __restrict__ icov /*[Nk][Z][9]*/,
    const double* __restrict__ lw /*[3][Nmu]*/, const double* __restrict__ shot, const int* __restrict__ ch_zb,
    const int* __restrict__ ch_off, const int* __restrict__ ch_cnt, const int* __restrict__ mem_s,
    double* __restrict__ partial) {
  __shared__ SDev sp;
  __shared__ RowC rows[CHI_ROWS];
  __shared__ double slw[3][CHI_ROWS];
  __shared__ double red[LEG_CH][TILE / 32];
  __shared__ double mb[LEG_CH], mx[LEG_CH];
  __shared__ int ms[LEG_CH];
  __shared__ MathTab mt;
  const int tid = threadIdx.x;
  const int zb = ch_zb[blockIdx.y], off = ch_off[blockIdx.y], cnt = ch_cnt[blockIdx.y];
  const int s0 = mem_s[off];
  if (tid < (int)(sizeof(SDev) / sizeof(double)))
    reinterpret_cast<double*>(&sp)[tid] = reinterpret_cast<const double*>(P.sdev + (size_t)zb * P.S + s0)[tid];
  math_tables_fill(mt, tid, TILE);
  __syncthreads();
  if (tid < LEG_CH) {
    const int sj = mem_s[off + min(tid, cnt - 1)];
    ms[tid] = sj;
    mb[tid] = P.sdev[(size_t)zb * P.S + sj].bias;
    mx[tid] = 1.0 / P.nbar[zb] + shot[(size_t)sj * P.Z + zb] + sp.pshot;
  }
  for (int r = tid; r < P.Nmu; r += TILE) {
    rows[r] = row_consts(sp, P.flags, __ldg(P.mu + r), __ldg(P.smu + r));
    slw[0][r] = lw[r], slw[1][r] = lw[P.Nmu + r], slw[2][r] = lw[2 * P.Nmu + r];
  }
  __syncthreads();
  const int nkb = (P.Nk + TILE - 1) / TILE;
  double acc[LEG_CH];
#pragma unroll
  for (int j = 0; j < LEG_CH; ++j) acc[j] = 0.0;
  for (int kb = blockIdx.x; kb < nkb; kb += gridDim.x) {
    const int ki = kb * TILE + tid;
    if (ki >= P.Nk) continue;
    const double k = __ldg(P.k + ki);
    int hA = -1, hB = -1, hN = -1;
    double a0[LEG_CH], a2[LEG_CH], a4[LEG_CH];
#pragma unroll
    for (int j = 0; j < LEG_CH; ++j) a0[j] = a2[j] = a4[j] = 0.0;
    for (int r = 0; r < P.Nmu; ++r) {
      double err2, qf, F, W;
      eval_x(mt, sp, rows[r], P.flags, k, hA, hB, hN, err2, qf, F, W);  // (eval_x_plain measured no faster here)
      const double We = W * tab_exp_neg(mt, -err2);
      const double w0 = slw[0][r], w2 = slw[1][r], w4 = slw[2][r];
#pragma unroll
      for (int j = 0; j < LEG_CH; ++j)
        if (j < cnt) {
          const double kaiser = fma(mb[j], qf, F);
          const double Pt = fma(kaiser * kaiser, We, mx[j]);
          a0[j] = fma(w0, Pt, a0[j]);
          a2[j] = fma(w2, Pt, a2[j]);
          a4[j] = fma(w4, Pt, a4[j]);
        }
    }
    const double* pd = pl_data + ((size_t)ki * P.Z + zb) * 3;
    const double* M = icov + ((size_t)ki * P.Z + zb) * 9;
    const double p0 = __ldg(pd), p2 = __ldg(pd + 1), p4 = __ldg(pd + 2);
    double m[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) m[e] = __ldg(M + e);
#pragma unroll
    for (int j = 0; j < LEG_CH; ++j)
      if (j < cnt) {
        const double d0 = p0 - a0[j], d1 = p2 - a2[j], d2 = p4 - a4[j];
        acc[j] += d0 * (m[0] * d0 + m[1] * d1 + m[2] * d2) + d1 * (m[3] * d0 + m[4] * d1 + m[5] * d2) +
                  d2 * (m[6] * d0 + m[7] * d1 + m[8] * d2);
      }
  }
#pragma unroll
  for (int j = 0; j < LEG_CH; ++j) {
    double a = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0) red[j][tid >> 5] = a;
  }
  __syncthreads();
  if (tid < cnt) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < TILE / 32; ++w) t += red[tid][w];
    partial[((size_t)ms[tid] * P.Z + zb) * gridDim.x + blockIdx.x] = t;
  }
}

// ------------------------------------------------------------------ likelihood of a large group through its moments
// Members of a group differ only in the galaxy bias b and in the additive noise m, and the theory is a polynomial in
// them: P_t,j - P_t,0 = (b_j^2 - b_0^2) a + (b_j - b_0) c + (m_j - m_0) with a = qf^2 We, c = 2 qf F We per cell and
// member 0 the group's reference.  Both likelihoods are quadratic in P_t, so chi2_j = v^T M v with
// v = (b_j^2 - b_0^2, b_j - b_0, m_j - m_0, 1) and a 4x4 matrix M summed over the lattice ONCE per group and bin --
// a nuisance-parameter batch of thousands of points costs one lattice pass per bin instead of one per 16 (wedges) or
// 4 (multipoles) points.  The residual of the reference member is formed exactly as the direct kernels form it, so
// a member equal to the reference (v = (0,0,0,1)) gets the direct kernel's value; the others differ from it by
// rounding only (tests/test_like_gpu.py compares the two paths).
// mom[(group * gridDim.x + cta) * MOM_N + e]: coefficients of v_i v_j, i <= j, in row-major order of (i, j).
constexpr int MOM_N = 10;
constexpr int MOM_MIN = 4;  // smaller groups take the direct kernels

template <int N>
__device__ __forceinline__ void block_sum_store(double (&acc)[N], double (*red)[TILE / 32], double* __restrict__ out) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double a = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0) red[j][tid >> 5] = a;
  }
  __syncthreads();
  if (tid < N) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < TILE / 32; ++w) t += red[tid][w];
    out[tid] = t;
  }
}

__global__ void __launch_bounds__(TILE, 3) spectro_wedge_moments_kernel(SpectroParams P, const double* __restrict__ data,
                                                                        const double* __restrict__ shot,
                                                                        const int* __restrict__ g_zb,
                                                                        const int* __restrict__ g_off,
                                                                        const int* __restrict__ mem_s,
                                                                        double* __restrict__ mom) {
  __shared__ SDev sp;
  __shared__ RowC rows[CHI_ROWS];
  __shared__ double wrow[CHI_ROWS];
  __shared__ double red[MOM_N][TILE / 32];
  __shared__ MathTab mt;
  const int tid = threadIdx.x;
  const int zb = g_zb[blockIdx.y], s0 = mem_s[g_off[blockIdx.y]];
  if (tid < (int)(sizeof(SDev) / sizeof(double)))
    reinterpret_cast<double*>(&sp)[tid] = reinterpret_cast<const double*>(P.sdev + (size_t)zb * P.S + s0)[tid];
  math_tables_fill(mt, tid, TILE);
  __syncthreads();
  const double b0 = sp.bias, m0 = 1.0 / P.nbar[zb] + shot[(size_t)s0 * P.Z + zb] + sp.pshot;
  const int64_t lattice = (int64_t)P.Nmu * P.Nk;
  const double pref = 2.0 * P.vol[zb] * INV_8PI2;  // spectro_like.py:162-178
  const double* __restrict__ dz = data + (size_t)zb * lattice;
  const int nkb = (P.Nk + TILE - 1) / TILE;
  const bool plain = (sp.flags & SD_PLAIN) != 0;
  double acc[MOM_N];
#pragma unroll
  for (int j = 0; j < MOM_N; ++j) acc[j] = 0.0;
  for (int r0 = 0; r0 < P.Nmu; r0 += CHI_ROWS) {
    const int nr = min(CHI_ROWS, P.Nmu - r0);
    __syncthreads();
    for (int r = tid; r < nr; r += TILE) {
      rows[r] = row_consts(sp, P.flags, __ldg(P.mu + r0 + r), __ldg(P.smu + r0 + r));
      wrow[r] = __ldg(P.wmu + r0 + r) * pref;
    }
    __syncthreads();
    for (int kb = blockIdx.x; kb < nkb; kb += gridDim.x) {
      const int ki = kb * TILE + tid;
      if (ki >= P.Nk) continue;
      const double k = __ldg(P.k + ki);
      const double wcol = __ldg(P.wk + ki) * (k * k);
      const double* __restrict__ dcol = dz + (size_t)r0 * P.Nk + ki;
      int hA = -1, hB = -1, hN = -1;
      for (int r = 0; r < nr; ++r) {
        const double Pd = __ldg(dcol + (size_t)r * P.Nk);
        double err2, qf, F, W;
        if (plain && hA >= 0) {
          qf = 1.0;
          eval_x_plain(mt, sp, rows[r], k, hA, hN, err2, F, W);
        } else {
          eval_x(mt, sp, rows[r], P.flags, k, hA, hB, hN, err2, qf, F, W);
        }
        const double We = W * tab_exp_neg(mt, -err2);
        const double rPd = fast_rcp(Pd);
        const double w = wrow[r] * wcol;
        const double kaiser = fma(b0, qf, F);
        const double Pt = fma(kaiser * kaiser, We, m0);
        const double g3 = (Pd - Pt) * rPd;  // the reference member's term, as spectro_chi2_kernel forms it
        const double qw = qf * We;
        const double g0 = -(qf * qw) * rPd, g1 = -(2.0 * (F * qw)) * rPd, g2 = -rPd;
        const double w0 = w * g0, w1 = w * g1, w2 = w * g2;
        acc[0] = fma(w0, g0, acc[0]);
        acc[1] = fma(w0, g1, acc[1]);
        acc[2] = fma(w0, g2, acc[2]);
        acc[3] = fma(w0, g3, acc[3]);
        acc[4] = fma(w1, g1, acc[4]);
        acc[5] = fma(w1, g2, acc[5]);
        acc[6] = fma(w1, g3, acc[6]);
        acc[7] = fma(w2, g2, acc[7]);
        acc[8] = fma(w2, g3, acc[8]);
        acc[9] = fma(w * g3, g3, acc[9]);
      }
    }
  }
  // coefficients of v_i v_j: the off-diagonal ones count twice
  acc[1] *= 2.0, acc[2] *= 2.0, acc[3] *= 2.0, acc[5] *= 2.0, acc[6] *= 2.0, acc[8] *= 2.0;
  block_sum_store(acc, red, mom + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * MOM_N);
}

// The same for the Legendre multipoles (see spectro_legendre_chi2_kernel): a thread owns a wavenumber, accumulates
// over the mu rows the multipoles of a, c and of the reference member's spectrum, and adds G^T C^-1 G of its (k, bin)
// with G = [-a_l, -c_l, -sum_mu lw_l, P_l,data - P_l,0] (3 x 4).
__global__ void __launch_bounds__(TILE, 3) spectro_legendre_moments_kernel(
    SpectroParams P, const double* __restrict__ pl_data /*[Nk][Z][3]*/, const double* __restrict__ icov /*[Nk][Z][9]*/,
    const double* __restrict__ lw /*[3][Nmu]*/, const double* __restrict__ shot, const int* __restrict__ g_zb,
    const int* __restrict__ g_off, const int* __restrict__ mem_s, double* __restrict__ mom) {
  __shared__ SDev sp;
  __shared__ RowC rows[CHI_ROWS];
  __shared__ double slw[3][CHI_ROWS];
  __shared__ double red[MOM_N][TILE / 32];
  __shared__ double lwsum[3];
  __shared__ MathTab mt;
  const int tid = threadIdx.x;
  const int zb = g_zb[blockIdx.y], s0 = mem_s[g_off[blockIdx.y]];
  if (tid < (int)(sizeof(SDev) / sizeof(double)))
    reinterpret_cast<double*>(&sp)[tid] = reinterpret_cast<const double*>(P.sdev + (size_t)zb * P.S + s0)[tid];
  math_tables_fill(mt, tid, TILE);
  __syncthreads();
  for (int r = tid; r < P.Nmu; r += TILE) {
    rows[r] = row_consts(sp, P.flags, __ldg(P.mu + r), __ldg(P.smu + r));
#pragma unroll
    for (int l = 0; l < 3; ++l) slw[l][r] = __ldg(lw + (size_t)l * P.Nmu + r);
  }
  __syncthreads();
  if (tid < 3) {  // multipoles of the additive noise: m sum_mu lw_l(mu), summed in row order as the direct kernel does
    double t = 0.0;
    for (int r = 0; r < P.Nmu; ++r) t += slw[tid][r];
    lwsum[tid] = t;
  }
  __syncthreads();
  const double b0 = sp.bias, m0 = 1.0 / P.nbar[zb] + shot[(size_t)s0 * P.Z + zb] + sp.pshot;
  const bool plain = (sp.flags & SD_PLAIN) != 0;
  const int nkb = (P.Nk + TILE - 1) / TILE;
  double acc[MOM_N];
#pragma unroll
  for (int j = 0; j < MOM_N; ++j) acc[j] = 0.0;
  for (int kb = blockIdx.x; kb < nkb; kb += gridDim.x) {
    const int ki = kb * TILE + tid;
    if (ki >= P.Nk) continue;
    const double k = __ldg(P.k + ki);
    int hA = -1, hB = -1, hN = -1;
    double A[3] = {0.0, 0.0, 0.0}, C[3] = {0.0, 0.0, 0.0}, R[3] = {0.0, 0.0, 0.0};
    for (int r = 0; r < P.Nmu; ++r) {
      double err2, qf, F, W;
      if (plain && hA >= 0) {
        qf = 1.0;
        eval_x_plain(mt, sp, rows[r], k, hA, hN, err2, F, W);
      } else {
        eval_x(mt, sp, rows[r], P.flags, k, hA, hB, hN, err2, qf, F, W);
      }
      const double We = W * tab_exp_neg(mt, -err2);
      const double kaiser = fma(b0, qf, F);
      const double Pt = fma(kaiser * kaiser, We, m0);  // the reference member, as spectro_legendre_chi2_kernel forms it
      const double qw = qf * We;
      const double a = qf * qw, c = 2.0 * (F * qw);
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double wl = slw[l][r];
        A[l] = fma(wl, a, A[l]);
        C[l] = fma(wl, c, C[l]);
        R[l] = fma(wl, Pt, R[l]);
      }
    }
    const double* pd = pl_data + ((size_t)ki * P.Z + zb) * 3;
    const double* M = icov + ((size_t)ki * P.Z + zb) * 9;
    double m[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) m[e] = __ldg(M + e);
    double g[4][3], t[4][3];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      g[0][l] = -A[l];
      g[1][l] = -C[l];
      g[2][l] = -lwsum[l];
      g[3][l] = __ldg(pd + l) - R[l];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)  // t_j = C^-1 g_j
#pragma unroll
      for (int l = 0; l < 3; ++l) t[j][l] = m[l * 3 + 0] * g[j][0] + m[l * 3 + 1] * g[j][1] + m[l * 3 + 2] * g[j][2];
    int e = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i; j < 4; ++j) {
        double q = g[i][0] * t[j][0] + g[i][1] * t[j][1] + g[i][2] * t[j][2];
        if (j != i) q += g[j][0] * t[i][0] + g[j][1] * t[i][1] + g[j][2] * t[i][2];
        acc[e++] += q;
      }
  }
  block_sum_store(acc, red, mom + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * MOM_N);
}

// chi2 of every member of the moment groups: block = group; the group's moments are summed over its CTAs in order,
// then a thread per member evaluates v^T M v.  Written where the direct kernels write (slot 0 of the member's
// (point, bin) partials, zeros in the other slots), so that one reduction serves both paths.
__global__ void __launch_bounds__(128) spectro_moments_apply_kernel(SpectroParams P, const double* __restrict__ shot,
                                                                    const int* __restrict__ g_zb,
                                                                    const int* __restrict__ g_off,
                                                                    const int* __restrict__ g_cnt,
                                                                    const int* __restrict__ mem_s,
                                                                    const double* __restrict__ mom, int gx_mom,
                                                                    double* __restrict__ partial, int gx_part) {
  __shared__ double c[MOM_N];
  const int g = blockIdx.x, zb = g_zb[g], off = g_off[g], cnt = g_cnt[g];
  if (threadIdx.x < MOM_N) {
    double t = 0.0;
    for (int cta = 0; cta < gx_mom; ++cta) t += mom[((size_t)g * gx_mom + cta) * MOM_N + threadIdx.x];
    c[threadIdx.x] = t;
  }
  __syncthreads();
  const int s0 = mem_s[off];
  const SDev* sd = P.sdev + (size_t)zb * P.S;
  const double b0 = sd[s0].bias, base = 1.0 / P.nbar[zb], ps = sd[s0].pshot;
  const double m0 = base + shot[(size_t)s0 * P.Z + zb] + ps;
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
    const int s = mem_s[off + j];
    const double b = sd[s].bias, mj = base + shot[(size_t)s * P.Z + zb] + ps;
    const double v0 = (b - b0) * (b + b0), v1 = b - b0, v2 = mj - m0;
    const double chi = c[9] + v0 * (c[3] + v0 * c[0] + v1 * c[1] + v2 * c[2]) + v1 * (c[6] + v1 * c[4] + v2 * c[5]) +
                       v2 * (c[8] + v2 * c[7]);
    double* out = partial + ((size_t)s * P.Z + zb) * gx_part;
    out[0] = chi;
    for (int q = 1; q < gx_part; ++q) out[q] = 0.0;
  }
}

// stand-alone wedge chi^2 on given lattices: grid (gx, Z, S)
__global__ void __launch_bounds__(TILE) wedge_chi2_kernel(int Z, int Nmu, int Nk, const double* __restrict__ theory,
                                                          const double* __restrict__ data, const double* __restrict__ k,
                                                          const double* __restrict__ wk, const double* __restrict__ wmu,
                                                          const double* __restrict__ vol, double* __restrict__ partial) {
  __shared__ double red[TILE / 32];
  const int zb = blockIdx.y, s = blockIdx.z, tid = threadIdx.x;
  const int64_t lattice = (int64_t)Nmu * Nk;
  const double pref = 2.0 * vol[zb] * INV_8PI2;
  double acc = 0.0;
  for (int64_t cell = (int64_t)blockIdx.x * TILE + tid; cell < lattice; cell += (int64_t)gridDim.x * TILE) {
    const int mi = (int)(cell / Nk), ki = (int)(cell - (int64_t)mi * Nk);
    const double kk = k[ki];
    const double Pd = data[(size_t)zb * lattice + cell];
    const double Pt = theory[((size_t)s * Z + zb) * lattice + cell];
    const double r = (Pd - Pt) / Pd;
    acc += (wmu[mi] * wk[ki]) * (pref * kk * kk) * (r * r);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < TILE / 32; ++w) t += red[w];
    partial[((size_t)s * Z + zb) * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void spectro_chi2_reduce_kernel(int S, int per, const double* __restrict__ partial, double* __restrict__ chi2) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double t = 0.0;
  for (int i = 0; i < per; ++i) t += partial[(size_t)s * per + i];
  chi2[s] = t;
}

}  // namespace

// ------------------------------------------------------------------------------ plan
struct cfp_spectro_plan {
  cfp_ctx* ctx = nullptr;
  const cfp_bank* bank = nullptr;
  int S = 0, Z = 0, Nmu = 0, Nk = 0, npar = 0, nb = 0, nnw = 0, NC = 0;
  unsigned flags = 0;
  int ps_src = 0, tab_pl = 0, tab_f = 0;
  int maxint = 0, maxncy = 0, grid_x = 0;
  int64_t cell_begin = 0, cell_end = 0;
  double *k_d = nullptr, *mu_d = nullptr, *smu_d = nullptr, *wk_d = nullptr, *wmu_d = nullptr;
  double *scal_d = nullptr, *pnw_k_d = nullptr, *pnw_p_d = nullptr, *nbar_d = nullptr, *vol_d = nullptr;
  double *st_w_d = nullptr, *st_den_d = nullptr, *st_rden_d = nullptr, *zbin_d = nullptr;
  double kmin = 0.0, kmax = 0.0;
  int *alias_d = nullptr, *ps_bin_d = nullptr, *st_ptr_d = nullptr, *st_par_d = nullptr, *dead_d = nullptr;
  double *blk_d = nullptr, *blknw_d = nullptr, *dcoef_d = nullptr;
  SDev* sdev_d = nullptr;
  int *work_d = nullptr, *nwork_d = nullptr, *shared_d = nullptr;
  double *partial_d = nullptr, *fisher_d = nullptr;
  double *lnp_d = nullptr, *derivs_d = nullptr, *veff_d = nullptr;
  PairPlan* pairs_d = nullptr;     // [Z] plan of the pair kernel per bin (written by the setup block of every run)
  double* partial2_d = nullptr;    // [Z][grid2_x][80] partials of the pair kernel
  double2* wd_d = nullptr;         // [Z][rows of the slice][Nk] {weight, shot-noise derivative} of the weights pre-pass
  int grid2_x = 0, wd_rows = 0;
  int pair_state = 0;              // bins the pair kernel takes when nothing is materialised: 0 unknown, 1 all, 2 none, 3 some
  int last_mat = -1;
  bool mid_recorded = false;       // kevm was recorded by the last run (the pair kernel ran)
  bool ran = false;
  cudaStream_t last_st = nullptr;  // stream of the most recent run (the pool releases on the context stream)
  std::vector<double> scal_h;      // host copy of the scalar blocks (the likelihood groups points by them)
  cudaEvent_t kev0 = nullptr, kev1 = nullptr, kevm = nullptr;  // lattice pass begin / end, and between pre-pass and main kernel
  // begin / end of this plan's most recent run or chi2 call, on the stream it ran on: fetch waits on run1 (plans of
  // one context may run concurrently on different streams, a context-wide event would be the wrong one)
  cudaEvent_t run0 = nullptr, run1 = nullptr;
  std::vector<void*> owned;
};

static void spectro_free(cfp_spectro_plan* p) {
  if (p->run0) cudaEventDestroy(p->run0);
  if (p->run1) cudaEventDestroy(p->run1);
  if (p->kev0) cudaEventDestroy(p->kev0);
  if (p->kev1) cudaEventDestroy(p->kev1);
  if (p->kevm) cudaEventDestroy(p->kevm);
  for (void* q : p->owned) cfp_dev_free(p->ctx, q);
  cfp_dev_free(p->ctx, p->lnp_d);
  cfp_dev_free(p->ctx, p->derivs_d);
  cfp_dev_free(p->ctx, p->veff_d);
  cfp_dev_free(p->ctx, p->partial_d);
  cfp_dev_free(p->ctx, p->partial2_d);
  cfp_dev_free(p->ctx, p->wd_d);
  delete p;
}

#define SP_UP(T, src, cnt, dst)                                \
  do {                                                         \
    int _rc = cfp_upload<T>(ctx, src, cnt, &(dst));            \
    if (_rc != CFP_OK) {                                       \
      spectro_free(pl);                                        \
      return _rc;                                              \
    }                                                          \
    pl->owned.push_back((void*)(dst));                         \
  } while (0)

extern "C" int cfp_spectro_plan_create(cfp_ctx* ctx, const cfp_bank* bank, int S, int Z, int Nmu, int Nk,
                                       int npar, const double* k_h, const double* mu_h, const double* wk_h,
                                       const double* wmu_h, const double* scal_h, const int32_t* alias_h,
                                       int nnw, const double* pnw_k_h, const double* pnw_p_h,
                                       const double* stw_h, const double* stden_h, const int32_t* ps_bin_h,
                                       int ps_src, const double* nbar_h, const double* vol_h, unsigned flags,
                                       int tab_plin, int tab_f, cfp_spectro_plan** out) {
  CFP_REQUIRE(ctx && bank && out, "NULL ctx/bank/out");
  CFP_REQUIRE(k_h && mu_h && wk_h && wmu_h && scal_h && alias_h && pnw_k_h && pnw_p_h && stw_h && stden_h &&
                  ps_bin_h && nbar_h && vol_h,
              "NULL input array");
  CFP_REQUIRE(S >= 1 && Z >= 1 && Nmu >= 1 && Nk >= 1 && npar >= 1 && nnw >= 2, "bad dimensions");
  CFP_REQUIRE(npar <= MAX_NB * 8, "npar exceeds 48");
  CFP_REQUIRE(ps_src >= 0 && ps_src < S, "ps_src out of range");
  CFP_REQUIRE(tab_plin >= 0 && tab_plin < bank->n_tables && tab_f >= 0 && tab_f < bank->n_tables, "bad table slot");
  *out = nullptr;
  CFP_CUDA(cudaSetDevice(ctx->device));
  const int NC = bank->n_cosmo;
  for (int s = 0; s < S; ++s)
    for (int z = 0; z < Z; ++z) {
      const double cd = scal_h[((size_t)s * Z + z) * CFP_SP_NSCAL + CFP_SP_COSMO];
      CFP_REQUIRE(cd >= 0 && cd < NC && cd == std::floor(cd), "cosmology index out of range");
      const int a = alias_h[s * Z + z];
      CFP_REQUIRE(a == s || a == 0, "alias must be s or 0");
    }
  cfp_spectro_plan* pl = new cfp_spectro_plan();
  pl->ctx = ctx;
  pl->bank = bank;
  pl->S = S, pl->Z = Z, pl->Nmu = Nmu, pl->Nk = Nk, pl->npar = npar, pl->nnw = nnw, pl->NC = NC;
  pl->nb = (npar + 7) / 8;
  pl->flags = flags;
  pl->ps_src = ps_src;
  pl->tab_pl = tab_plin, pl->tab_f = tab_f;
  pl->cell_begin = 0;
  pl->cell_end = (int64_t)Nmu * Nk;

  // knot windows
  int maxint = 1, maxncy = 1;
  for (int c = 0; c < NC; ++c)
    for (int t = 0; t < 2; ++t) {
      const int64_t* d = bank->desc(c, t == 0 ? tab_plin : tab_f);
      if (d[0] < 8 || d[1] < 8) {
        cfp_set_error("cfp_spectro_plan_create: cosmology %d lacks table slot %d", c, t == 0 ? tab_plin : tab_f);
        spectro_free(pl);
        return CFP_ERR_INVALID;
      }
      maxint = std::max(maxint, (int)d[1] - 7);
      maxncy = std::max(maxncy, (int)d[1] - 4);
    }
  pl->maxint = maxint;
  pl->maxncy = maxncy;

  // stencil in CSR form per stencil point; dead (parameter, bin) pairs
  std::vector<int> st_ptr(S + 1, 0), st_par, dead((size_t)npar * Z, 0);
  std::vector<double> st_w;
  for (int s = 0; s < S; ++s) {
    for (int p = 0; p < npar; ++p)
      if (stw_h[(size_t)p * S + s] != 0.0 && ps_bin_h[p] < 0) {
        st_par.push_back(p);
        st_w.push_back(stw_h[(size_t)p * S + s]);
      }
    st_ptr[s + 1] = (int)st_par.size();
  }
  for (int p = 0; p < npar; ++p)
    for (int z = 0; z < Z; ++z) {
      if (ps_bin_h[p] >= 0) continue;
      bool all_fid = true;
      for (int s = 0; s < S; ++s)
        if (stw_h[(size_t)p * S + s] != 0.0 && !(s == 0 || alias_h[s * Z + z] == 0)) all_fid = false;
      dead[(size_t)p * Z + z] = all_fid ? 1 : 0;
    }
  if (st_par.empty()) {
    st_par.push_back(0);
    st_w.push_back(0.0);
  }
  std::vector<double> smu(Nmu), zbin(Z);
  for (int i = 0; i < Nmu; ++i) smu[i] = std::sqrt(1.0 - mu_h[i] * mu_h[i]);
  for (int z = 0; z < Z; ++z) zbin[z] = scal_h[(size_t)z * CFP_SP_NSCAL + CFP_SP_Z];

  SP_UP(double, k_h, (size_t)Nk, pl->k_d);
  SP_UP(double, mu_h, (size_t)Nmu, pl->mu_d);
  SP_UP(double, smu.data(), (size_t)Nmu, pl->smu_d);
  SP_UP(double, wk_h, (size_t)Nk, pl->wk_d);
  SP_UP(double, wmu_h, (size_t)Nmu, pl->wmu_d);
  SP_UP(double, scal_h, (size_t)S * Z * CFP_SP_NSCAL, pl->scal_d);
  pl->scal_h.assign(scal_h, scal_h + (size_t)S * Z * CFP_SP_NSCAL);
  SP_UP(int, alias_h, (size_t)S * Z, pl->alias_d);
  SP_UP(double, pnw_k_h, (size_t)NC * nnw, pl->pnw_k_d);
  SP_UP(double, pnw_p_h, (size_t)NC * Z * nnw, pl->pnw_p_d);
  SP_UP(double, stden_h, (size_t)npar, pl->st_den_d);
  std::vector<double> rden(npar);
  for (int p = 0; p < npar; ++p) rden[p] = 1.0 / stden_h[p];
  SP_UP(double, rden.data(), (size_t)npar, pl->st_rden_d);
  pl->kmin = *std::min_element(k_h, k_h + Nk);
  pl->kmax = *std::max_element(k_h, k_h + Nk);
  SP_UP(int, ps_bin_h, (size_t)npar, pl->ps_bin_d);
  SP_UP(double, nbar_h, (size_t)Z, pl->nbar_d);
  SP_UP(double, vol_h, (size_t)Z, pl->vol_d);
  SP_UP(int, st_ptr.data(), st_ptr.size(), pl->st_ptr_d);
  SP_UP(int, st_par.data(), st_par.size(), pl->st_par_d);
  SP_UP(double, st_w.data(), st_w.size(), pl->st_w_d);
  SP_UP(int, dead.data(), dead.size(), pl->dead_d);
  SP_UP(double, zbin.data(), (size_t)Z, pl->zbin_d);
  SP_UP(double, nullptr, (size_t)NC * Z * 2 * maxint * CUBIC_BLK, pl->blk_d);
  SP_UP(double, nullptr, (size_t)NC * Z * (nnw - 1) * LIN_BLK, pl->blknw_d);
  SP_UP(SDev, nullptr, (size_t)Z * S, pl->sdev_d);
  SP_UP(int, nullptr, (size_t)Z * S, pl->work_d);
  SP_UP(int, nullptr, (size_t)Z * 8, pl->nwork_d);
  SP_UP(int, nullptr, (size_t)NC * Z, pl->shared_d);
  SP_UP(double, nullptr, (size_t)NC * Z * 2 * maxncy, pl->dcoef_d);
  SP_UP(double, nullptr, (size_t)Z * npar * npar, pl->fisher_d);
  SP_UP(PairPlan, nullptr, (size_t)Z, pl->pairs_d);
  // the uploads read caller-owned host memory asynchronously: finish them before returning
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    cfp_set_error("cfp_spectro_plan_create: upload failed");
    spectro_free(pl);
    return CFP_ERR_CUDA;
  }
  *out = pl;
  return CFP_OK;
}

extern "C" int cfp_spectro_plan_destroy(cfp_spectro_plan* plan) {
  if (!plan) return CFP_OK;
  cudaSetDevice(plan->ctx->device);
  // wait for THIS plan's last run on the caller's stream (its own event), not for the whole stream: in a pipeline of
  // batches the stream already holds the next plan's kernels
  if (plan->last_st && plan->last_st != plan->ctx->stream) {
    if (plan->run1)
      cudaEventSynchronize(plan->run1);
    else
      cudaStreamSynchronize(plan->last_st);
  }
  spectro_free(plan);
  return CFP_OK;
}

extern "C" int cfp_spectro_plan_set_slice(cfp_spectro_plan* plan, int64_t cell_begin, int64_t cell_end) {
  CFP_REQUIRE(plan != nullptr, "plan is NULL");
  CFP_REQUIRE(cell_begin >= 0 && cell_begin <= cell_end && cell_end <= (int64_t)plan->Nmu * plan->Nk,
              "slice out of range");
  plan->cell_begin = cell_begin;
  plan->cell_end = cell_end;
  return CFP_OK;
}

template <int NB>
static cudaError_t launch_fisher(const SpectroParams& P, dim3 grid, size_t smem, int row_lo, int nrows, cudaStream_t st) {
  cudaError_t e =
      cudaFuncSetAttribute(spectro_fisher_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  spectro_fisher_kernel<NB><<<grid, TILE, smem, st>>>(P, row_lo, nrows);
  return cudaGetLastError();
}

static size_t fisher_smem(int nb, int S) {
  const int T = nb * (nb + 1) / 2;
  return std::max((size_t)(nb * 8 * (TILE + LDPAD) + TILE) * sizeof(double) + (size_t)S * (sizeof(SDev) + FISHER_ROWS * sizeof(RowC)) +
                      sizeof(MathTab),
                  (size_t)(TILE / 32) * T * 64 * sizeof(double));
}

static void spectro_fill_params(cfp_spectro_plan* pl, int materialize, SpectroParams& P) {
  P.S = pl->S, P.Z = pl->Z, P.Nmu = pl->Nmu, P.Nk = pl->Nk, P.npar = pl->npar, P.nb = pl->nb, P.nnw = pl->nnw;
  P.NC = pl->NC, P.flags = pl->flags, P.ps_src = pl->ps_src, P.materialize = materialize, P.maxint = pl->maxint;
  P.cell_begin = pl->cell_begin, P.cell_end = pl->cell_end;
  P.k = pl->k_d, P.mu = pl->mu_d, P.smu = pl->smu_d, P.wk = pl->wk_d, P.wmu = pl->wmu_d;
  P.scal = pl->scal_d, P.alias = pl->alias_d, P.st_ptr = pl->st_ptr_d, P.st_par = pl->st_par_d;
  P.st_w = pl->st_w_d, P.st_den = pl->st_den_d, P.st_rden = pl->st_rden_d, P.kmin = pl->kmin, P.kmax = pl->kmax, P.ps_bin = pl->ps_bin_d, P.dead = pl->dead_d;
  P.nbar = pl->nbar_d, P.vol = pl->vol_d, P.blob = pl->bank->blob_d, P.desc = pl->bank->desc_d;
  P.n_tables = pl->bank->n_tables, P.tab_pl = pl->tab_pl, P.tab_f = pl->tab_f;
  P.blk = pl->blk_d, P.blknw = pl->blknw_d, P.sdev = pl->sdev_d, P.work = pl->work_d, P.nwork = pl->nwork_d;
  P.shared_tmp = pl->shared_d;
  P.partial = pl->partial_d;
  P.lnp = pl->lnp_d, P.derivs = pl->derivs_d, P.veff = pl->veff_d;
  P.pairs = pl->pairs_d;
}

static int spectro_prepare(cfp_spectro_plan* pl, const SpectroParams& P, cudaStream_t st) {
  cfp_ctx* ctx = pl->ctx;
  PrepParams pp;
  pp.NC = pl->NC, pp.Z = pl->Z, pp.nnw = pl->nnw, pp.maxint = pl->maxint, pp.maxncy = pl->maxncy;
  pp.blob = pl->bank->blob_d, pp.desc = pl->bank->desc_d, pp.n_tables = pl->bank->n_tables;
  pp.tab[0] = pl->tab_pl, pp.tab[1] = pl->tab_f;
  pp.zbin = pl->zbin_d, pp.dcoef = pl->dcoef_d, pp.blk = pl->blk_d;
  pp.pnw_k = pl->pnw_k_d, pp.pnw_p = pl->pnw_p_d, pp.blknw = pl->blknw_d;
  const size_t cls_bytes = (P.materialize & 16) ? 0 : (size_t)pl->S * sizeof(int);
  CFP_REQUIRE(cls_bytes <= 40 * 1024, "too many stencil points for a Fisher work list (batch mode has no limit)");
  spectro_prepare_kernel<<<dim3(pl->NC * pl->Z, 4), 512, cls_bytes, st>>>(pp, P, pl->Z);
  ctx->launches++;
  CFP_CUDA(cudaGetLastError());
  return CFP_OK;
}

// Chunks of a likelihood batch: per bin, points whose scalar blocks are identical except for the bias are evaluated
// together (up to `chunk` per chunk).  Grouping through an open-addressing table keyed by a multiply-add hash of the
// block's words (bias excluded); a hit is verified against the group's first member, so unequal blocks never share a
// group whatever the hash does.  Groups come in the order of their first member, members in index order (the first
// member is the reference point of the moment kernels).  16384 points x 4 bins: 0.8 ms (sorting (hash, index) pairs
// with an FNV chain per block took 2.9 ms and was the largest part of a batch whose kernels take 0.3 ms).
// Chunk list = [groups of 2..chunk points | single points] (n_grp, n_one of each).
static void spectro_build_chunks(const cfp_spectro_plan* pl, int chunk, std::vector<int>& ch_zb, std::vector<int>& ch_off,
                                 std::vector<int>& ch_cnt, std::vector<int>& mem_s, int& n_grp, int& n_one) {
  const int S = pl->S, Z = pl->Z;
  constexpr int NS = CFP_SP_NSCAL, NB = CFP_SP_BIAS;
  std::vector<int> one_zb, one_off;
  const double* sc = pl->scal_h.data();
  uint64_t K[NS];
  {
    uint64_t x = 0x9E3779B97F4A7C15ull;  // splitmix64: one odd multiplier per word of the block
    for (int q = 0; q < NS; ++q) {
      x += 0x9E3779B97F4A7C15ull;
      uint64_t z = x;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      K[q] = (z ^ (z >> 31)) | 1ull;
    }
  }
  int T = 1;
  while (T < 2 * S) T <<= 1;
  std::vector<int> table((size_t)T), gid((size_t)S), rep, cnt, start, fill, members((size_t)S);
  std::vector<uint64_t> gh;
  for (int zb = 0; zb < Z; ++zb) {
    std::fill(table.begin(), table.end(), -1);
    rep.clear(), cnt.clear(), gh.clear();
    for (int sidx = 0; sidx < S; ++sidx) {
      const double* x = sc + ((size_t)sidx * Z + zb) * NS;
      uint64_t w[NS];
      std::memcpy(w, x, sizeof(w));
      uint64_t h0 = 0, h1 = 0;
      for (int q = 0; q < NB; ++q) h0 += w[q] * K[q];
      for (int q = NB + 1; q < NS; ++q) h1 += w[q] * K[q];
      uint64_t h = h0 + h1;
      h ^= h >> 29;
      int slot = (int)(h & (uint64_t)(T - 1)), g;
      for (;;) {
        g = table[slot];
        if (g < 0) break;
        if (gh[g] == h) {
          const double* y = sc + ((size_t)rep[g] * Z + zb) * NS;
          if (std::memcmp(x, y, NB * sizeof(double)) == 0 &&
              std::memcmp(x + NB + 1, y + NB + 1, (NS - NB - 1) * sizeof(double)) == 0)
            break;
        }
        slot = (slot + 1) & (T - 1);
      }
      if (g < 0) {
        g = (int)rep.size();
        table[slot] = g;
        rep.push_back(sidx), gh.push_back(h), cnt.push_back(0);
      }
      gid[sidx] = g;
      cnt[g]++;
    }
    const int ng = (int)rep.size();
    start.assign((size_t)ng + 1, 0);
    for (int g = 0; g < ng; ++g) start[g + 1] = start[g] + cnt[g];
    fill.assign(start.begin(), start.end() - 1);
    for (int sidx = 0; sidx < S; ++sidx) members[fill[gid[sidx]]++] = sidx;
    for (int g = 0; g < ng; ++g) {
      const int* same = members.data() + start[g];
      const int n = cnt[g];
      for (int o = 0; o < n; o += chunk) {
        const int c = std::min(chunk, n - o);
        auto& Z_ = c == 1 ? one_zb : ch_zb;
        auto& O_ = c == 1 ? one_off : ch_off;
        Z_.push_back(zb), O_.push_back((int)mem_s.size());
        if (c > 1) ch_cnt.push_back(c);
        mem_s.insert(mem_s.end(), same + o, same + o + c);
      }
    }
  }
  n_grp = (int)ch_zb.size(), n_one = (int)one_zb.size();
  ch_zb.insert(ch_zb.end(), one_zb.begin(), one_zb.end());
  ch_off.insert(ch_off.end(), one_off.begin(), one_off.end());
  ch_cnt.insert(ch_cnt.end(), (size_t)n_one, 1);
}

static int spectro_chi2_impl(cfp_spectro_plan* pl, const double* data_h, const double* data_dev, const double* shot_h,
                             const double* shot_data_h, double* chi2_h, void* stream);

extern "C" int cfp_spectro_plan_chi2(cfp_spectro_plan* pl, const double* data_h, const double* shot_h,
                                     const double* shot_data_h, double* chi2_h, void* stream) {
  return spectro_chi2_impl(pl, data_h, nullptr, shot_h, shot_data_h, chi2_h, stream);
}

// The groups of a batch as the likelihood entry points launch them: [groups of >= MOM_MIN members (moment kernels) |
// smaller groups (one chunk each for the direct kernels) | single points].  CFP_NO_MOMENTS=1 sends every group through
// the direct kernels in chunks of `chunk` (tests compare the two paths).
static void spectro_build_groups(const cfp_spectro_plan* pl, int chunk, std::vector<int>& zb, std::vector<int>& off,
                                 std::vector<int>& cnt, std::vector<int>& mem_s, int& n_mom, int& n_grp, int& n_one) {
  const char* env = std::getenv("CFP_NO_MOMENTS");
  if (env && env[0] == '1') {
    n_mom = 0;
    spectro_build_chunks(pl, chunk, zb, off, cnt, mem_s, n_grp, n_one);
    return;
  }
  std::vector<int> gz, go, gc;
  int ng = 0;
  spectro_build_chunks(pl, INT_MAX, gz, go, gc, mem_s, ng, n_one);
  n_mom = n_grp = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < ng; ++i)
      if ((gc[i] >= MOM_MIN) == (pass == 0)) {
        zb.push_back(gz[i]), off.push_back(go[i]), cnt.push_back(gc[i]);
        ++(pass == 0 ? n_mom : n_grp);
      }
  zb.insert(zb.end(), gz.begin() + ng, gz.end());
  off.insert(off.end(), go.begin() + ng, go.end());
  cnt.insert(cnt.end(), gc.begin() + ng, gc.end());
}

// CTAs per moment group: fill the GPU (3 resident CTAs per SM), at most one per block of wavenumbers
static int moments_grid_x(const cfp_ctx* ctx, int n_mom, int nkb) {
  const int64_t slots = 3 * (int64_t)ctx->sm_count;
  return (int)std::max<int64_t>(1, std::min<int64_t>(nkb, slots / std::max(n_mom, 1)));
}

extern "C" int cfp_spectro_plan_chi2_dev(cfp_spectro_plan* pl, const double* data_d, const double* shot_h,
                                         double* chi2_h, void* stream) {
  CFP_REQUIRE(data_d != nullptr, "data_d is NULL");
  return spectro_chi2_impl(pl, nullptr, data_d, shot_h, nullptr, chi2_h, stream);
}

static int spectro_chi2_impl(cfp_spectro_plan* pl, const double* data_h, const double* data_dev, const double* shot_h,
                             const double* shot_data_h, double* chi2_h, void* stream) {
  CFP_REQUIRE(pl && shot_h && chi2_h, "NULL argument");
  cfp_ctx* ctx = pl->ctx;
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const int S = pl->S, Z = pl->Z;
  const size_t lattice = (size_t)pl->Nmu * pl->Nk;
  const int64_t ntile = (int64_t)((lattice + TILE - 1) / TILE);
  std::vector<int> ch_zb, ch_off, ch_cnt, mem_s;
  int n_mom = 0, n_grp = 0, n_one = 0;
  spectro_build_groups(pl, CHI_CH, ch_zb, ch_off, ch_cnt, mem_s, n_mom, n_grp, n_one);
  const int nch = n_mom + n_grp + n_one;
  // CTAs per chunk: each takes every gx-th block of TILE wavenumbers.  Pick the split that needs the fewest
  // (waves of resident CTAs) x (blocks per CTA), i.e. fills the GPU for small batches and avoids a ragged last wave
  // for large ones.
  const int nkb = (pl->Nk + TILE - 1) / TILE;
  int gx = 1;
  {
    // resident CTAs: 3 per SM for the group kernel, 5 for the single-point kernel
    const int64_t slots = (n_grp >= n_one ? 3 : 5) * (int64_t)ctx->sm_count;
    int64_t best = INT64_MAX;
    for (int g = 1; g <= std::min(nkb, 64); ++g) {
      const int64_t ctas = (int64_t)g * std::max(n_grp, n_one);
      const int64_t cost = ((ctas + slots - 1) / slots) * ((nkb + g - 1) / g);
      if (cost < best) best = cost, gx = g;
    }
  }
  int *ch_zb_d = nullptr, *ch_off_d = nullptr, *ch_cnt_d = nullptr, *mem_s_d = nullptr;
  double *data_d = nullptr, *shot_d = nullptr, *sd_d = nullptr, *part_d = nullptr, *chi2_d = nullptr, *mom_d = nullptr;
  const int gxm = moments_grid_x(ctx, n_mom, nkb);
  auto cleanup = [&]() {
    for (void* q : {(void*)data_d, (void*)shot_d, (void*)sd_d, (void*)part_d, (void*)chi2_d, (void*)ch_zb_d, (void*)ch_off_d,
                    (void*)ch_cnt_d, (void*)mem_s_d, (void*)mom_d})
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
  if (!data_dev) CHI_TRY(cfp_upload<double>(ctx, data_h, (size_t)Z * lattice, &data_d));
  CHI_TRY(cfp_upload(ctx, shot_h, (size_t)S * Z, &shot_d));
  if (shot_data_h) CHI_TRY(cfp_upload(ctx, shot_data_h, (size_t)Z, &sd_d));
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)S * Z * gx, &part_d));
  CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)S, &chi2_d));
  if (n_mom) CHI_TRY(cfp_upload<double>(ctx, nullptr, (size_t)n_mom * gxm * MOM_N, &mom_d));
  CHI_TRY(cfp_upload(ctx, ch_zb.data(), ch_zb.size(), &ch_zb_d));
  CHI_TRY(cfp_upload(ctx, ch_off.data(), ch_off.size(), &ch_off_d));
  CHI_TRY(cfp_upload(ctx, ch_cnt.data(), ch_cnt.size(), &ch_cnt_d));
  CHI_TRY(cfp_upload(ctx, mem_s.data(), mem_s.size(), &mem_s_d));
  CFP_CUDA(cudaStreamSynchronize(ctx->stream));  // the chunk lists are host temporaries
  if (!pl->run0) {
    cudaEventCreate(&pl->run0);
    cudaEventCreate(&pl->run1);
  }
  cudaEventRecord(pl->run0, st);
  SpectroParams P;
  spectro_fill_params(pl, 16, P);
  CHI_TRY(spectro_prepare(pl, P, st));
  if (!data_h && !data_dev) {
    const int gd = (int)std::min<int64_t>(ntile, 2 * ctx->sm_count);
    spectro_data_kernel<<<dim3(gd, Z), TILE, 0, st>>>(P, sd_d, data_d);
    ctx->launches++;
  }
  const double* data_use = data_dev ? data_dev : data_d;
  for (int c0 = 0; c0 < n_mom; c0 += 65535) {
    const int nc = std::min(65535, n_mom - c0);
    spectro_wedge_moments_kernel<<<dim3(gxm, nc), TILE, 0, st>>>(P, data_use, shot_d, ch_zb_d + c0, ch_off_d + c0, mem_s_d,
                                                                mom_d + (size_t)c0 * gxm * MOM_N);
    ctx->launches++;
  }
  if (n_mom) {
    spectro_moments_apply_kernel<<<n_mom, 128, 0, st>>>(P, shot_d, ch_zb_d, ch_off_d, ch_cnt_d, mem_s_d, mom_d, gxm, part_d, gx);
    ctx->launches++;
  }
  for (int c0 = n_mom; c0 < n_mom + n_grp; c0 += 65535) {
    const int nc = std::min(65535, n_mom + n_grp - c0);
    spectro_chi2_kernel<CHI_CH, 3><<<dim3(gx, nc), TILE, 0, st>>>(P, data_use, shot_d, ch_zb_d + c0, ch_off_d + c0,
                                                                 ch_cnt_d + c0, mem_s_d, part_d);
    ctx->launches++;
  }
  for (int c0 = n_mom + n_grp; c0 < nch; c0 += 65535) {
    const int nc = std::min(65535, nch - c0);
    spectro_chi2_kernel<1, 5><<<dim3(gx, nc), TILE, 0, st>>>(P, data_use, shot_d, ch_zb_d + c0, ch_off_d + c0,
                                                            ch_cnt_d + c0, mem_s_d, part_d);
    ctx->launches++;
  }
  spectro_chi2_reduce_kernel<<<(S + 127) / 128, 128, 0, st>>>(S, Z * gx, part_d, chi2_d);
  ctx->launches++;
  cudaEventRecord(pl->run1, st);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(chi2_h, chi2_d, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  float ms = 0.f;
  if (e == cudaSuccess && cudaEventElapsedTime(&ms, pl->run0, pl->run1) == cudaSuccess) ctx->last_ms = ms;
  cleanup();
#undef CHI_TRY
  if (e != cudaSuccess) {
    cfp_set_error("cfp_spectro_plan_chi2: %s", cudaGetErrorString(e));
    return CFP_ERR_CUDA;
  }
  return CFP_OK;
}

extern "C" int cfp_spectro_plan_chi2_legendre(cfp_spectro_plan* pl, const double* pl_data_h, const double* icov_h,
                                              const double* lw_h, const double* shot_h, double* chi2_h, void* stream) {
  CFP_REQUIRE(pl && pl_data_h && icov_h && lw_h && shot_h && chi2_h, "NULL argument");
  CFP_REQUIRE(pl->Nmu <= CHI_ROWS, "the fused multipole path stages every mu row at once (Nmu <= 256)");
  cfp_ctx* ctx = pl->ctx;
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const int S = pl->S, Z = pl->Z, Nk = pl->Nk;
  std::vector<int> ch_zb, ch_off, ch_cnt, mem_s;
  int n_mom = 0, n_grp = 0, n_one = 0;
  spectro_build_groups(pl, LEG_CH, ch_zb, ch_off, ch_cnt, mem_s, n_mom, n_grp, n_one);
  const int nch = n_mom + n_grp + n_one, ndir = n_grp + n_one;
  const int nkb = (Nk + TILE - 1) / TILE;
  const int gxm = moments_grid_x(ctx, n_mom, nkb);
  int gx = 1;
  {
    const int64_t slots = 4 * (int64_t)ctx->sm_count;
    int64_t best = INT64_MAX;
    for (int g = 1; g <= std::min(nkb, 64); ++g) {
      const int64_t cost = (((int64_t)g * std::max(ndir, 1) + slots - 1) / slots) * ((nkb + g - 1) / g);
      if (cost < best) best = cost, gx = g;
    }
  }
  int *ch_zb_d = nullptr, *ch_off_d = nullptr, *ch_cnt_d = nullptr, *mem_s_d = nullptr;
  double *pd_d = nullptr, *ic_d = nullptr, *lw_d = nullptr, *shot_d = nullptr, *part_d = nullptr, *chi2_d = nullptr;
  double* mom_d = nullptr;
  auto cleanup = [&]() {
    for (void* q : {(void*)pd_d, (void*)ic_d, (void*)lw_d, (void*)shot_d, (void*)part_d, (void*)chi2_d, (void*)ch_zb_d,
                    (void*)ch_off_d, (void*)ch_cnt_d, (void*)mem_s_d, (void*)mom_d})
      cfp_dev_free(ctx, q);
  };
  int rc;
#define LEG_TRY(x)      \
  do {                  \
    rc = (x);           \
    if (rc != CFP_OK) { \
      cleanup();        \
      return rc;        \
    }                   \
  } while (0)
  LEG_TRY(cfp_upload(ctx, pl_data_h, (size_t)Nk * Z * 3, &pd_d));
  LEG_TRY(cfp_upload(ctx, icov_h, (size_t)Nk * Z * 9, &ic_d));
  LEG_TRY(cfp_upload(ctx, lw_h, (size_t)3 * pl->Nmu, &lw_d));
  LEG_TRY(cfp_upload(ctx, shot_h, (size_t)S * Z, &shot_d));
  LEG_TRY(cfp_upload<double>(ctx, nullptr, (size_t)S * Z * gx, &part_d));
  LEG_TRY(cfp_upload<double>(ctx, nullptr, (size_t)S, &chi2_d));
  if (n_mom) LEG_TRY(cfp_upload<double>(ctx, nullptr, (size_t)n_mom * gxm * MOM_N, &mom_d));
  LEG_TRY(cfp_upload(ctx, ch_zb.data(), ch_zb.size(), &ch_zb_d));
  LEG_TRY(cfp_upload(ctx, ch_off.data(), ch_off.size(), &ch_off_d));
  LEG_TRY(cfp_upload(ctx, ch_cnt.data(), ch_cnt.size(), &ch_cnt_d));
  LEG_TRY(cfp_upload(ctx, mem_s.data(), mem_s.size(), &mem_s_d));
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    cleanup();
    cfp_set_error("cfp_spectro_plan_chi2_legendre: upload failed");
    return CFP_ERR_CUDA;
  }
  if (!pl->run0) {
    cudaEventCreate(&pl->run0);
    cudaEventCreate(&pl->run1);
  }
  cudaEventRecord(pl->run0, st);
  SpectroParams P;
  spectro_fill_params(pl, 16, P);
  LEG_TRY(spectro_prepare(pl, P, st));
  for (int c0 = 0; c0 < n_mom; c0 += 65535) {
    const int nc = std::min(65535, n_mom - c0);
    spectro_legendre_moments_kernel<<<dim3(gxm, nc), TILE, 0, st>>>(P, pd_d, ic_d, lw_d, shot_d, ch_zb_d + c0, ch_off_d + c0,
                                                                   mem_s_d, mom_d + (size_t)c0 * gxm * MOM_N);
    ctx->launches++;
  }
  if (n_mom) {
    spectro_moments_apply_kernel<<<n_mom, 128, 0, st>>>(P, shot_d, ch_zb_d, ch_off_d, ch_cnt_d, mem_s_d, mom_d, gxm, part_d, gx);
    ctx->launches++;
  }
  for (int c0 = n_mom; c0 < nch; c0 += 65535) {
    const int nc = std::min(65535, nch - c0);
    spectro_legendre_chi2_kernel<<<dim3(gx, nc), TILE, 0, st>>>(P, pd_d, ic_d, lw_d, shot_d, ch_zb_d + c0, ch_off_d + c0,
                                                               ch_cnt_d + c0, mem_s_d, part_d);
    ctx->launches++;
  }
  spectro_chi2_reduce_kernel<<<(S + 127) / 128, 128, 0, st>>>(S, Z * gx, part_d, chi2_d);
  ctx->launches++;
  cudaEventRecord(pl->run1, st);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(chi2_h, chi2_d, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  float ms = 0.f;
  if (e == cudaSuccess && cudaEventElapsedTime(&ms, pl->run0, pl->run1) == cudaSuccess) ctx->last_ms = ms;
  cleanup();
#undef LEG_TRY
  if (e != cudaSuccess) {
    cfp_set_error("cfp_spectro_plan_chi2_legendre: %s", cudaGetErrorString(e));
    return CFP_ERR_CUDA;
  }
  return CFP_OK;
}

extern "C" int cfp_spectro_plan_run(cfp_spectro_plan* pl, int materialize, void* stream) {
  CFP_REQUIRE(pl != nullptr, "plan is NULL");
  cfp_ctx* ctx = pl->ctx;
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const size_t lattice = (size_t)pl->Nmu * pl->Nk;
  bool allocated = false;
  if ((materialize & CFP_MAT_LNP) && !pl->lnp_d) {
    CFP_CUDA(cfp_dev_alloc(ctx, (void**)&pl->lnp_d, (size_t)pl->S * pl->Z * lattice * sizeof(double)));
    allocated = true;
  }
  if ((materialize & CFP_MAT_DERIVS) && !pl->derivs_d) {
    CFP_CUDA(cfp_dev_alloc(ctx, (void**)&pl->derivs_d, (size_t)pl->npar * pl->Z * lattice * sizeof(double)));
    CFP_CUDA(cfp_dev_alloc(ctx, (void**)&pl->veff_d, (size_t)pl->Z * lattice * sizeof(double)));
    allocated = true;
  }
  // tiles = (blocks of TILE wavenumbers) x (mu rows the slice touches)
  const int row_lo = (int)(pl->cell_begin / pl->Nk);
  const int nrows = std::max(1, (int)((pl->cell_end + pl->Nk - 1) / pl->Nk) - row_lo);
  const int64_t ntile = (int64_t)((pl->Nk + TILE - 1) / TILE) * nrows;
  const int T = pl->nb * (pl->nb + 1) / 2;
  // which kernel: the pair kernel takes the bins of a plain 3PT forecast (decided on the device by the setup block
  // of each run; the host learns the answer at the first fetch and then skips the launches that would do nothing)
  static const bool no_pairs = std::getenv("CFP_NO_PAIRS") != nullptr;
  const bool try_pairs = materialize == 0 && !no_pairs && pl->pair_state != 2;
  const bool run_general = !(try_pairs && pl->pair_state == 1);
  // persistent grid: one wave of (resident CTAs per SM) x (SM count) CTAs, never more CTAs than tiles
  const int minb = fisher_min_blocks(pl->nb);
  int gx = (int)std::min<int64_t>(ntile, std::max(1, (ctx->sm_count * minb + pl->Z - 1) / pl->Z));
  if (run_general && gx != pl->grid_x) {
    if (pl->partial_d && pl->last_st && pl->last_st != ctx->stream) CFP_CUDA(cudaStreamSynchronize(pl->last_st));
    cfp_dev_free(ctx, pl->partial_d);
    pl->partial_d = nullptr;
    CFP_CUDA(cfp_dev_alloc(ctx, (void**)&pl->partial_d, (size_t)pl->Z * gx * T * 64 * sizeof(double)));
    pl->grid_x = gx;
    allocated = true;
  }
  const int nquad = (pl->Nk + 3) / 4;
  static const int pk_warps = std::getenv("CFP_PK_WARPS") ? std::atoi(std::getenv("CFP_PK_WARPS")) : 16;
  const int gx2 = std::max(1, std::min((nquad + pk_warps - 1) / pk_warps, ctx->sm_count / pl->Z));
  if (try_pairs && (gx2 != pl->grid2_x || nrows != pl->wd_rows)) {
    if (pl->partial2_d && pl->last_st && pl->last_st != ctx->stream) CFP_CUDA(cudaStreamSynchronize(pl->last_st));
    cfp_dev_free(ctx, pl->partial2_d);
    cfp_dev_free(ctx, pl->wd_d);
    pl->partial2_d = nullptr, pl->wd_d = nullptr;
    CFP_CUDA(cfp_dev_alloc(ctx, (void**)&pl->partial2_d, (size_t)pl->Z * gx2 * 80 * sizeof(double)));
    CFP_CUDA(cfp_dev_alloc(ctx, (void**)&pl->wd_d, (size_t)pl->Z * nrows * (4 * nquad) * sizeof(double2)));
    pl->grid2_x = gx2;
    pl->wd_rows = nrows;
    allocated = true;
  }
  // pool allocations are ordered on the context stream; a caller's stream must not start before them
  if (allocated && st != ctx->stream) CFP_CUDA(cudaStreamSynchronize(ctx->stream));
  pl->last_st = st;
  pl->last_mat = materialize;
  if (!pl->run0) {
    CFP_CUDA(cudaEventCreate(&pl->run0));
    CFP_CUDA(cudaEventCreate(&pl->run1));
  }
  CFP_CUDA(cudaEventRecord(pl->run0, st));
  SpectroParams P;
  spectro_fill_params(pl, try_pairs ? 0 : (materialize | (no_pairs || pl->pair_state == 2 ? 32 : 0)), P);
  {
    int rc = spectro_prepare(pl, P, st);
    if (rc != CFP_OK) return rc;
  }
  if (!pl->kev0) {
    CFP_CUDA(cudaEventCreate(&pl->kev0));
    CFP_CUDA(cudaEventCreate(&pl->kev1));
    CFP_CUDA(cudaEventCreate(&pl->kevm));
  }
  CFP_CUDA(cudaEventRecord(pl->kev0, st));
  pl->mid_recorded = false;
  cudaError_t e = cudaSuccess;
  if (try_pairs) {
    // rows staged at once, and how finely the rows are cut into tasks: enough tasks for an even deal to the warps
    const int srows = std::min(nrows, PK_MAXROWS);
    const int nwarps = gx2 * pk_warps;
    int nseg = 1;
    double best = 0.0;
    for (int c = 1; c <= 4 && c <= nrows; ++c) {
      const int64_t tasks = (int64_t)nquad * c;
      const double eff = (double)tasks / (double)(((tasks + nwarps - 1) / nwarps) * nwarps);
      if (eff > best + 0.02) best = eff, nseg = c;
    }
    {
      // row segments: as many as keep the grid within ONE wave of resident CTAs (3 per SM), at most WG_ROWS rows each
      const int kb = (pl->Nk + TILE - 1) / TILE;
      const int wsegs = std::min(nrows, std::max((nrows + WG_ROWS - 1) / WG_ROWS, (ctx->sm_count * 3) / (kb * pl->Z)));
      const int wlen = (nrows + wsegs - 1) / wsegs;
      spectro_weights_kernel<<<dim3((pl->Nk + TILE - 1) / TILE, pl->Z, wsegs), TILE, 0, st>>>(P, row_lo, nrows, wlen,
                                                                                            pl->wd_d);
      cudaEventRecord(pl->kevm, st);
      pl->mid_recorded = true;
    }
    ctx->launches++;
    const size_t smem2 = std::max((size_t)srows * 16 * sizeof(RowP), (size_t)pk_warps * 80 * sizeof(double)) + sizeof(MathTab) +
                         16 * sizeof(SDev) + sizeof(PairPlan);
#define CFP_LAUNCH_PK(WV)                                                                                             \
  case WV:                                                                                                            \
    e = cudaFuncSetAttribute(spectro_fisher_pairs_kernel<WV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2); \
    if (e == cudaSuccess)                                                                                             \
      spectro_fisher_pairs_kernel<WV><<<dim3(gx2, pl->Z), WV * 32, smem2, st>>>(P, row_lo, nrows, srows, nseg,         \
                                                                               pl->wd_d, pl->partial2_d);             \
    break;
    switch (pk_warps) {
      CFP_LAUNCH_PK(12)  // (A/B variants, CFP_PK_WARPS: 16 warps at 128 registers is the measured optimum)
      CFP_LAUNCH_PK(20)
      default:
        e = cudaFuncSetAttribute(spectro_fisher_pairs_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e == cudaSuccess)
          spectro_fisher_pairs_kernel<16><<<dim3(gx2, pl->Z), 16 * 32, smem2, st>>>(P, row_lo, nrows, srows, nseg, pl->wd_d,
                                                                                    pl->partial2_d);
        break;
    }
#undef CFP_LAUNCH_PK
    ctx->launches++;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
      cfp_set_error("spectro_fisher_pairs_kernel launch: %s", cudaGetErrorString(e));
      return CFP_ERR_CUDA;
    }
  }
  if (run_general) {
    const dim3 grid(gx, pl->Z);
    const size_t smem = fisher_smem(pl->nb, pl->S);
#define CFP_LAUNCH_NB(NBV)                                   \
  case NBV:                                                  \
    e = launch_fisher<NBV>(P, grid, smem, row_lo, nrows, st); \
    break;
    switch (pl->nb) {
      CFP_LAUNCH_NB(1)
      CFP_LAUNCH_NB(2)
      CFP_LAUNCH_NB(3)
      CFP_LAUNCH_NB(4)
      CFP_LAUNCH_NB(5)
      default:
        e = launch_fisher<6>(P, grid, smem, row_lo, nrows, st);
        break;
    }
#undef CFP_LAUNCH_NB
    ctx->launches++;
    if (e != cudaSuccess) {
      cfp_set_error("spectro_fisher_kernel launch: %s", cudaGetErrorString(e));
      return CFP_ERR_CUDA;
    }
  }
  CFP_CUDA(cudaEventRecord(pl->kev1, st));
  if (run_general) {
    spectro_reduce_kernel<<<dim3(pl->Z, T * 2), 1024, 0, st>>>(pl->partial_d, gx, pl->nb, pl->npar, pl->pairs_d, pl->fisher_d);
    ctx->launches++;
  }
  if (try_pairs) {
    spectro_reduce_pairs_kernel<<<dim3(pl->Z, 8), 1024, 0, st>>>(pl->partial2_d, gx2, pl->npar, pl->pairs_d, pl->fisher_d);
    ctx->launches++;
  }
  CFP_CUDA(cudaGetLastError());
  CFP_CUDA(cudaEventRecord(pl->run1, st));
  pl->ran = true;
  return CFP_OK;
}

extern "C" double cfp_spectro_plan_kernel_ms(cfp_spectro_plan* pl) {
  if (!pl || !pl->ran || !pl->kev1) return -1.0;
  cudaSetDevice(pl->ctx->device);
  if (cudaEventSynchronize(pl->kev1) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, pl->kev0, pl->kev1) != cudaSuccess) return -1.0;
  return ms;
}

extern "C" int cfp_spectro_plan_terms(cfp_spectro_plan* pl, int s, double* terms_h) {
  CFP_REQUIRE(pl && terms_h, "NULL argument");
  CFP_REQUIRE(s >= 0 && s < pl->S, "stencil point out of range");
  cfp_ctx* ctx = pl->ctx;
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)5 * pl->Z * pl->Nmu * pl->Nk;
  double* out_d = nullptr;
  CFP_CUDA(cfp_dev_alloc(ctx, (void**)&out_d, n * sizeof(double)));
  SpectroParams P;
  spectro_fill_params(pl, 16, P);  // (batch mode: scalar blocks and piece tables, no work list)
  int rc = spectro_prepare(pl, P, st);
  if (rc == CFP_OK) {
    const int gx = std::max(1, std::min((int)(((size_t)pl->Nmu * pl->Nk + TILE - 1) / TILE), (ctx->sm_count * 3) / pl->Z + 1));
    spectro_terms_kernel<<<dim3(gx, pl->Z), TILE, 0, st>>>(P, s, out_d);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(terms_h, out_d, n * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      cfp_set_error("cfp_spectro_plan_terms: %s", cudaGetErrorString(e));
      rc = CFP_ERR_CUDA;
    }
  }
  cfp_dev_free(ctx, out_d);
  return rc;
}

extern "C" int cfp_math_selftest(cfp_ctx* ctx, int which, const double* x_h, int n, double* y_h) {
  CFP_REQUIRE(ctx && x_h && y_h && n >= 1 && which >= 0 && which <= 5, "bad argument");
  CFP_CUDA(cudaSetDevice(ctx->device));
  double *x_d = nullptr, *y_d = nullptr;
  int rc = cfp_upload<double>(ctx, x_h, (size_t)n, &x_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, nullptr, (size_t)n, &y_d);
  if (rc == CFP_OK) {
    math_selftest_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(which, n, x_d, y_d);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(y_h, y_d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      cfp_set_error("cfp_math_selftest: %s", cudaGetErrorString(e));
      rc = CFP_ERR_CUDA;
    }
  }
  cfp_dev_free(ctx, x_d);
  cfp_dev_free(ctx, y_d);
  return rc;
}

extern "C" double cfp_spectro_plan_kernel_ms_part(cfp_spectro_plan* pl, int which) {
  if (which == 0) return cfp_spectro_plan_kernel_ms(pl);
  if (!pl || !pl->ran || !pl->kev1 || !pl->mid_recorded || which < 0 || which > 2) return -1.0;
  cudaSetDevice(pl->ctx->device);
  if (cudaEventSynchronize(pl->kev1) != cudaSuccess) return -1.0;
  float ms = 0.f;
  const cudaError_t e = which == 1 ? cudaEventElapsedTime(&ms, pl->kev0, pl->kevm) : cudaEventElapsedTime(&ms, pl->kevm, pl->kev1);
  return e == cudaSuccess ? (double)ms : -1.0;
}

extern "C" const double* cfp_spectro_plan_fisher_d(const cfp_spectro_plan* plan) {
  return plan ? plan->fisher_d : nullptr;
}

extern "C" int cfp_spectro_plan_fetch(cfp_spectro_plan* pl, double* fisher_h, double* lnp_h, double* derivs_h,
                                      double* veff_h) {
  CFP_REQUIRE(pl != nullptr, "plan is NULL");
  if (!pl->ran) {
    cfp_set_error("cfp_spectro_plan_fetch: plan has not been run");
    return CFP_ERR_STATE;
  }
  cfp_ctx* ctx = pl->ctx;
  CFP_CUDA(cudaSetDevice(ctx->device));
  CFP_CUDA(cudaEventSynchronize(pl->run1));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, pl->run0, pl->run1) == cudaSuccess) ctx->last_ms = ms;
  if (pl->last_mat == 0 && pl->pair_state == 0) {
    // which bins the pair kernel took (the setup block decided): later runs skip the launch that would do nothing
    std::vector<PairPlan> pp(pl->Z);
    CFP_CUDA(cudaMemcpy(pp.data(), pl->pairs_d, (size_t)pl->Z * sizeof(PairPlan), cudaMemcpyDeviceToHost));
    int n_ok = 0;
    for (const PairPlan& q : pp) n_ok += q.ok ? 1 : 0;
    pl->pair_state = n_ok == pl->Z ? 1 : (n_ok == 0 ? 2 : 3);
  }
  const size_t lattice = (size_t)pl->Nmu * pl->Nk;
  if (fisher_h)
    CFP_CUDA(cudaMemcpy(fisher_h, pl->fisher_d, (size_t)pl->Z * pl->npar * pl->npar * sizeof(double), cudaMemcpyDeviceToHost));
  if (lnp_h) {
    CFP_REQUIRE(pl->lnp_d != nullptr, "lnP was not materialised by the last run");
    CFP_CUDA(cudaMemcpy(lnp_h, pl->lnp_d, (size_t)pl->S * pl->Z * lattice * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (derivs_h) {
    CFP_REQUIRE(pl->derivs_d != nullptr, "derivatives were not materialised by the last run");
    CFP_CUDA(cudaMemcpy(derivs_h, pl->derivs_d, (size_t)pl->npar * pl->Z * lattice * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (veff_h) {
    CFP_REQUIRE(pl->veff_d != nullptr, "veff was not materialised by the last run");
    CFP_CUDA(cudaMemcpy(veff_h, pl->veff_d, (size_t)pl->Z * lattice * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return CFP_OK;
}

extern "C" int cfp_wedge_chi2(cfp_ctx* ctx, int S, int Z, int Nmu, int Nk, const double* theory_h, const double* data_h,
                              const double* k_h, const double* wk_h, const double* wmu_h, const double* vol_h,
                              double* chi2_h) {
  CFP_REQUIRE(ctx && theory_h && data_h && k_h && wk_h && wmu_h && vol_h && chi2_h, "NULL argument");
  CFP_REQUIRE(S >= 1 && S <= 65535 && Z >= 1 && Nmu >= 1 && Nk >= 1, "bad dimensions");
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t lattice = (size_t)Nmu * Nk;
  const int gx = (int)std::max<size_t>(1, std::min<size_t>((lattice + TILE - 1) / TILE, 64));
  double *th_d = nullptr, *da_d = nullptr, *k_d = nullptr, *wk_d = nullptr, *wmu_d = nullptr, *vol_d = nullptr;
  double *part_d = nullptr, *chi2_d = nullptr;
  auto cleanup = [&]() {
    for (void* q : {(void*)th_d, (void*)da_d, (void*)k_d, (void*)wk_d, (void*)wmu_d, (void*)vol_d, (void*)part_d,
                    (void*)chi2_d})
      cfp_dev_free(ctx, q);
  };
  int rc = cfp_upload(ctx, theory_h, (size_t)S * Z * lattice, &th_d);
  if (rc == CFP_OK) rc = cfp_upload(ctx, data_h, (size_t)Z * lattice, &da_d);
  if (rc == CFP_OK) rc = cfp_upload(ctx, k_h, (size_t)Nk, &k_d);
  if (rc == CFP_OK) rc = cfp_upload(ctx, wk_h, (size_t)Nk, &wk_d);
  if (rc == CFP_OK) rc = cfp_upload(ctx, wmu_h, (size_t)Nmu, &wmu_d);
  if (rc == CFP_OK) rc = cfp_upload(ctx, vol_h, (size_t)Z, &vol_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, nullptr, (size_t)S * Z * gx, &part_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, nullptr, (size_t)S, &chi2_d);
  if (rc != CFP_OK) {
    cleanup();
    return rc;
  }
  wedge_chi2_kernel<<<dim3(gx, Z, S), TILE, 0, st>>>(Z, Nmu, Nk, th_d, da_d, k_d, wk_d, wmu_d, vol_d, part_d);
  spectro_chi2_reduce_kernel<<<(S + 127) / 128, 128, 0, st>>>(S, Z * gx, part_d, chi2_d);
  ctx->launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(chi2_h, chi2_d, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) {
    cfp_set_error("cfp_wedge_chi2: %s", cudaGetErrorString(e));
    return CFP_ERR_CUDA;
  }
  return CFP_OK;
}
