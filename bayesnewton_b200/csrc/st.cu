// Dense spatio-temporal path: Kalman filter, RTS smoother and the
// projection steps either side of them for SpatioTemporalKernel (bayesnewton/kernels.py:385-586):
//     state x = [u_1; ...; u_M], u_i the n-dimensional temporal state of spatial inducing point i,
//     A_k = I_M (x) A_t(dt_k),  Pinf = I_M (x) Pinf_t,  H = I_M (x) H_t   (d = M n, D = M).
// The state covariance is a dense d x d matrix (d = 512 at C4), so one time step is a Cholesky
// factorisation plus a few d^3 contractions; time stays sequential (the reference's lax.scan,
// ops.py:154-180, 288-311) and the parallelism is INSIDE the step:
//
//   * one persistent kernel per pass, one CTA per SM, the whole time loop inside the kernel; the
//     phases of a step are separated by a grid barrier (an L2 atomic counter), not by launches;
//   * the Kronecker structure is never multiplied out: predict is a per-(n x n)-block rotation;
//   * every solve of the step rides on ONE blocked Cholesky: the matrices that need L^-T applied
//     from the right (P^- H^T, the residual, an identity for L^-1, ...) are stacked under the SPD
//     matrix and swept by the same left-looking panels (32 columns per phase), so their rows are
//     extra parallelism for the panel phase instead of extra triangular solves after it;
//   * the remaining work is C = A B^T tiles (NT form only), fp64 FMA from shared-memory tiles.
//
// The batched projection steps (compute_full_pseudo_lik, basemodels.py:676-687, and the Gaussian
// KL term) use the same panel code with one CTA per time step.
#include "common.cuh"
#include "gen.cuh"
#include "condstats.cuh"

namespace bn {
namespace st {

constexpr int NB = 32;     // panel width = row-block height
constexpr int NTH = 256;   // threads per CTA
constexpr int BK = 32;     // k-slab of the tile product

__host__ __device__ inline int pad32(int x) { return (x + 31) / 32 * 32; }

__device__ __forceinline__ double ldg(const double* p) { return __ldcg(p); }  // L2: data other CTAs wrote in this kernel

// phase profile of the last filter / smoother launch (measurement aid, bn_st_profile): SM cycles summed over steps
//  0 assemble  1 barrier after assemble  2 Cholesky sweep (CTA 0: all of it)  3 tile phases  4 their barriers
//  5 panel solve, 6 look-ahead, 7 barrier wait (CTA owning the LAST row block)  8 factor+invert+publish (sum over owners)
__device__ long long g_prof[16];
#define ST_PROF(slot, t0) do { if (threadIdx.x == 0) { long long t1__ = clock64(); atomicAdd((unsigned long long*)&g_prof[slot], (unsigned long long)(t1__ - (t0))); (t0) = t1__; } } while (0)

struct Smem {
    double a[64 * 36];     // A slab: [row][k], pitch 36 (conflict-free DMMA fragment loads)
    double b[64 * 36];     // B slab
    double L[NB][NB + 1];  // diagonal block / its Cholesky factor
    double Li[NB][NB + 1]; // inverse of the factor (lower)
    double X[NB][NB + 1];  // right-hand block before the triangular solve
    double XO[NB][NB + 1]; // solved block (kept for the diagonal update)
    double XA[NB][NB + 1]; // previous panel's column block of this row block
    double XB[NB][NB + 1]; // previous panel's column block of the diagonal row block
    double col[2][NB];     // column buffer of the in-warp factorisation
    double dinv[NB];       // reciprocal diagonal of the factor
};
constexpr size_t kSmemBytes = sizeof(Smem);  // ~84 KB: dynamic shared memory, one CTA per SM

// ---- grid barrier ---------------------------------------------------------------------------------
// All CTAs are co-resident (grid <= number of SMs, checked on the host).  The counter only grows; the
// host zeroes it before the launch.
struct GridSync {
    unsigned long long* ctr;   // ctr[0]: all CTAs;  ctr[16] (its own 128-byte line): the CTAs of a panel sweep
    unsigned long long epoch;
    unsigned long long epoch2 = 0ULL;
    __device__ void sync() {
        __syncthreads();
        if (threadIdx.x == 0) {
            epoch += gridDim.x;
            __threadfence();
            atomicAdd(ctr, 1ULL);
            while (*((volatile unsigned long long*)ctr) < epoch) { }
            __threadfence();
        }
        __syncthreads();
    }
    // barrier among the n CTAs that take part in a panel sweep: the other CTAs wait at the next full barrier, on a
    // different line, so their polling does not sit on the line the sweep synchronises through
    __device__ void sync_sub(int n) {
        __syncthreads();
        if (threadIdx.x == 0) {
            epoch2 += (unsigned long long)n;
            __threadfence();
            atomicAdd(ctr + 16, 1ULL);
            while (*((volatile unsigned long long*)(ctr + 16)) < epoch2) { }
            __threadfence();
        }
        __syncthreads();
    }
};

// ---- C tile = A B^T on the fp64 tensor pipe ---------------------------------------------------------------------
// acc (TM x TN outputs per CTA, TM, TN in {32, 64}) = sum_{k<K} A[row][k] * (SCALE ? s[k] : 1) * B[col][k]; rows >=
// arows / brows and k >= K read as zero.  8 warps as 2 x 4; a warp owns (TM/2) x (TN/4) outputs as m8n8k4 DMMA tiles
// (mma.sync.aligned.m8n8k4.row.col.f64: lane (g, t) = (lane / 4, lane % 4) feeds A[g][t] and B^T[g][t], and holds
// C[g][2t], C[g][2t+1]).  Operands are staged [row][k] with pitch 36 doubles, so the 32 lanes of a fragment load hit
// 16 distinct bank pairs twice -- the 2-wavefront minimum for 256 bytes; one fragment pair feeds 256 FMAs.  The next
// k-slab is fetched from L2 into registers while the current one is multiplied.
// acc[r][c] belongs to tile row trow<TM>(r), tile column tcol<TN>(c).
constexpr int kPitch = 36;

template <int TM>
__device__ __forceinline__ int trow(int r) {
    return ((threadIdx.x >> 7) & 1) * (TM / 2) + 8 * r + ((threadIdx.x & 31) >> 2);
}
template <int TN>
__device__ __forceinline__ int tcol(int c) {
    return ((threadIdx.x >> 5) & 3) * (TN / 4) + 8 * (c >> 1) + 2 * (threadIdx.x & 3) + (c & 1);
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int TM, int TN, bool SCALE>
__device__ __forceinline__ void tile_nt(const double* __restrict__ A, int lda, int arows, const double* __restrict__ B,
                                        int ldb, int brows, int K, int i0, int j0, const double* __restrict__ s,
                                        double (&acc)[TM / 16][TN / 16], Smem& sm) {
    constexpr int RM = TM / 16, RN = TN / 16, NI = TN / 32;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wm0 = ((threadIdx.x >> 7) & 1) * (TM / 2), wn0 = ((threadIdx.x >> 5) & 3) * (TN / 4);
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int c = 0; c < RN; ++c) acc[r][c] = 0.0;
    // staging: thread (ty, tx) moves k = tx, tx + 16 of rows ty + 16 r
    double pa[2][RM], pb[2][RN];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int kk = k0 + tx + 16 * h;
            const double sc = (SCALE && kk < K) ? ldg(s + kk) : 1.0;
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                const int row = i0 + ty + 16 * r;
                pa[h][r] = (row < arows && kk < K) ? ldg(A + (size_t)row * lda + kk) * sc : 0.0;
            }
#pragma unroll
            for (int r = 0; r < RN; ++r) {
                const int row = j0 + ty + 16 * r;
                pb[h][r] = (row < brows && kk < K) ? ldg(B + (size_t)row * ldb + kk) : 0.0;
            }
        }
    };
    if (K > 0) fetch(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int r = 0; r < RM; ++r) sm.a[(ty + 16 * r) * kPitch + tx + 16 * h] = pa[h][r];
#pragma unroll
            for (int r = 0; r < RN; ++r) sm.b[(ty + 16 * r) * kPitch + tx + 16 * h] = pb[h][r];
        }
        __syncthreads();
        if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[RM], bf[NI];
#pragma unroll
            for (int r = 0; r < RM; ++r) af[r] = sm.a[(wm0 + 8 * r + g) * kPitch + kk + t];
#pragma unroll
            for (int c = 0; c < NI; ++c) bf[c] = sm.b[(wn0 + 8 * c + g) * kPitch + kk + t];
#pragma unroll
            for (int r = 0; r < RM; ++r)
#pragma unroll
                for (int c = 0; c < NI; ++c) dmma(acc[r][2 * c], acc[r][2 * c + 1], af[r], bf[c]);
        }
    }
}

// ---- stacked blocked Cholesky ---------------------------------------------------------------------------
// T: rows x n (row-major, ld = n, both multiples of 32).  Rows [0,n) hold an SPD matrix S (lower triangle
// used), the rows below hold stacked right-hand sides Wstack.  After panels 0..n/32-1:
//     T[0:n] lower triangle = L (S = L L^T),   T[n:] = Wstack L^-T.

// 1 / sqrt(x) to fp64 rounding: the hardware's fp64 approximation (MUFU.RSQ64H, 2^-22 relative) + two Newton steps
// (error -> 1e-13 -> < 1 ulp); negative / NaN input gives NaN (a non-PD block poisons like cho_factor), 0 gives inf
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return y;
}

// sm.L (a 32 x 32 SPD block, lower triangle read) -> its Cholesky factor in sm.L (upper part zeroed), 1 / L_rr in
// sm.dinv.  Warp 0, lane r owns row r in registers.  Right-looking: once column c is scaled, the update of entry
// (c+1, c+1) -- the next pivot -- is the first of the rank-one updates, so the serial chain per pivot is
// shuffle -> rsqrt -> multiply -> one shared-memory round trip -> one FMA; the other updates fill the pipe behind it.
__device__ __forceinline__ void factor_warp(Smem& sm) {
    const int r = threadIdx.x;
    double a[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) a[q] = sm.L[r][q];
    double dg = sm.L[r][r];  // the lane's own diagonal entry, kept current from its own column values (no round trip)
    double dinv = 0.0;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        const double piv = __shfl_sync(0xffffffffu, dg, c);
        // 1 / sqrt(piv): hardware seed, one Newton step, the second one folded into the scaling of the column
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(piv));
        const double hx = 0.5 * piv;
        y = fma(y, fma(-hx * y, y, 0.5), y);
        const double e2 = fma(-hx * y, y, 0.5);
        const double ay = ((r == c) ? piv : a[c]) * y;
        double l = fma(ay, e2, ay);
        l = (r >= c) ? l : 0.0;
        if (r == c) dinv = fma(y, e2, y);
        dg = fma(-l, l, dg);
        a[c] = l;
        sm.col[c & 1][r] = l;
        __syncwarp();
#pragma unroll
        for (int q = c + 1; q < NB; ++q) a[q] = fma(-l, sm.col[c & 1][q], a[q]);
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) sm.L[r][q] = (q <= r) ? a[q] : 0.0;
    sm.dinv[r] = dinv;
}

// sm.L (a 32 x 32 SPD block) -> its Cholesky factor in sm.L (upper part zeroed) and the inverse factor in sm.Li.
// Executed by warp 0; a non-positive pivot gives NaN (cho_factor).  Callers sync before and after.
// Lane r keeps row r of L in registers (fully unrolled), so a column costs one broadcast shared-memory load per
// term, one shuffle for the pivot and one rsqrt; the inverse reuses the reciprocal pivots (no divisions).
__device__ __forceinline__ void factor_invert_warp(Smem& sm) {
    const int r = threadIdx.x;  // lane r owns row r
    double lr[NB];
    double dinv = 0.0;          // 1 / L[r][r]
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        double s0 = sm.L[r][c], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int q = 0; q < c; ++q) {
            const double lcq = sm.L[c][q];
            if ((q & 3) == 0) s0 = fma(-lr[q], lcq, s0);
            else if ((q & 3) == 1) s1 = fma(-lr[q], lcq, s1);
            else if ((q & 3) == 2) s2 = fma(-lr[q], lcq, s2);
            else s3 = fma(-lr[q], lcq, s3);
        }
        const double sv = (s0 + s1) + (s2 + s3);
        const double piv = __shfl_sync(0xffffffffu, sv, c);
        const double inv = fast_rsqrt(piv);  // NaN for piv < 0
        const double v = (r == c) ? piv * inv : (r > c ? sv * inv : 0.0);
        lr[c] = v;
        sm.L[r][c] = v;
        if (r == c) dinv = inv;
        __syncwarp();
    }
    // Li = L^-1 (lower): lane cc solves L x = e_cc by forward substitution
    const int cc = threadIdx.x;
    double x[NB];
#pragma unroll
    for (int rr = 0; rr < NB; ++rr) {
        double s0 = (rr == cc) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int q = 0; q < rr; ++q) {
            const double l = sm.L[rr][q];
            if ((q & 3) == 0) s0 = fma(-l, x[q], s0);
            else if ((q & 3) == 1) s1 = fma(-l, x[q], s1);
            else if ((q & 3) == 2) s2 = fma(-l, x[q], s2);
            else s3 = fma(-l, x[q], s3);
        }
        x[rr] = ((s0 + s1) + (s2 + s3)) * __shfl_sync(0xffffffffu, dinv, rr);
    }
#pragma unroll
    for (int rr = 0; rr < NB; ++rr) sm.Li[rr][cc] = x[rr];
}

// ---- single-CTA form (batched problems: one CTA per matrix), left-looking --------------------------------
// D_j = S_jj - sum_{p<j} L_jp L_jp^T  ->  sm.L = chol(D_j), sm.Li = its inverse.  All threads call.
__device__ void panel_factor(const double* T, int ld, int j, Smem& sm) {
    double acc[2][2];
    const double* rowj = T + (size_t)j * NB * ld;
    tile_nt<32, 32, false>(rowj, ld, NB, rowj, ld, NB, j * NB, 0, 0, nullptr, acc, sm);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int rr = trow<32>(r), cc = tcol<32>(c);
            sm.L[rr][cc] = ldg(rowj + (size_t)rr * ld + j * NB + cc) - acc[r][c];
        }
    __syncthreads();
    if (threadIdx.x < 32) factor_invert_warp(sm);
    __syncthreads();
}

// row block i of panel j: T[i, j] <- (T[i, j] - sum_{p<j} T[i, p] L_jp^T) L_jj^-T   (i != j);  T[j, j] <- L_jj
// k0: the first k0 columns of row block i are structurally zero (identity stacks) and are skipped in the k-loop
__device__ void panel_apply(double* T, int ld, int j, int i, Smem& sm, int k0 = 0) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double* out = T + (size_t)i * NB * ld + j * NB;
    if (i == j) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) out[(size_t)(ty + 16 * r) * ld + tx + 16 * c] = sm.L[ty + 16 * r][tx + 16 * c];
        return;
    }
    double acc[2][2];
    tile_nt<32, 32, false>(T + (size_t)i * NB * ld + k0, ld, NB, T + (size_t)j * NB * ld + k0, ld, NB, j * NB - k0, 0, 0, nullptr,
                           acc, sm);
    __syncthreads();  // sm.X may still be read by the previous row block's solve
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int rr = trow<32>(r), cc = tcol<32>(c);
            sm.X[rr][cc] = ldg(out + (size_t)rr * ld + cc) - acc[r][c];
        }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int rr = ty + 16 * r, cc = tx + 16 * c;
            double s = 0.0;
            for (int q = 0; q <= cc; ++q) s = fma(sm.X[rr][q], sm.Li[cc][q], s);
            out[(size_t)rr * ld + cc] = s;
        }
}

// Row blocks [id_lo, id_lo + n/32) may hold an identity (the stack that turns into L^-T): block id_lo + b is zero left
// of column 32 b, so it joins the sweep at panel b and its k-loops start there -- a third of the stacked work of an
// [S ; Y ; I] sweep disappears.
__device__ void chol_stack_cta(double* T, int n, int rows, Smem& sm, int id_lo = -1) {
    const int nsq = n / NB, ntot = rows / NB;
    for (int j = 0; j < nsq; ++j) {
        panel_factor(T, n, j, sm);
        for (int i = j; i < ntot; ++i) {
            int k0 = 0;
            if (id_lo >= 0 && i >= id_lo && i < id_lo + nsq) {
                const int b = i - id_lo;
                if (b > j) continue;  // still exactly zero in this block column
                k0 = b * NB;
            }
            panel_apply(T, n, j, i, sm, k0);
        }
        __syncthreads();
        __threadfence_block();
    }
}

// ---- grid form: row blocks owned by CTAs, one grid barrier per panel, look-ahead ---------------------------
// The sequential chain of a blocked Cholesky is: column block j of row block j+1 -> diagonal block j+1 -> its
// factor.  Everything else is kept off that chain:
//   * each row block has an owner CTA; in phase j the owner forms its block of column j from the part
//     accumulated ahead of time (Pacc, all panels p < j-1) plus ONE rank-32 term (panel j-1), and solves it
//     against the published inverse factor of diagonal block j;
//   * owners keep their own diagonal block up to date (right-looking, rank-32 per phase), so the owner of row
//     block j+1 can factor and invert it at once and publish L_{j+1}^-1 (double-buffered) before the barrier;
//   * meanwhile the other owners accumulate Pacc for phase j+1 (the long k-loop), overlapping the factorisation.
constexpr int kLpub = NB * NB + NB;  // a published diagonal factor: L (32 x 32) and 1 / diag L
struct ChSys {
    double* T;      // [rows, n]
    int n, rows;
    double* Pacc;   // [rows/32][32*32]  look-ahead accumulators
    double* Lpub;   // [2][kLpub]        published diagonal factors (double-buffered)
};

__device__ __forceinline__ void publish_factor(const ChSys& s, int jb, Smem& sm) {
    // sm.L holds the fully updated diagonal block jb: factor it, write L to T[jb, jb] and (L, 1 / diag L) to Lpub[jb & 1]
    __syncthreads();
    long long tp = clock64();
    if (threadIdx.x < 32) factor_warp(sm);
    __syncthreads();
    ST_PROF(8, tp);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double* out = s.T + (size_t)jb * NB * s.n + jb * NB;
    double* lp = s.Lpub + (size_t)(jb & 1) * kLpub;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int rr = ty + 16 * r, cc = tx + 16 * c;
            out[(size_t)rr * s.n + cc] = sm.L[rr][cc];
            lp[rr * NB + cc] = sm.L[rr][cc];
        }
    if (threadIdx.x < NB) lp[NB * NB + threadIdx.x] = sm.dinv[threadIdx.x];
}

__device__ void chol_sys_phase(const ChSys& s, int j, int lb, int LG, Smem& sm) {
    const int nsq = s.n / NB, ntot = s.rows / NB, ld = s.n;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    // first owned row block > j:  i = lb (mod LG)
    int i0 = lb;
    if (i0 <= j) i0 += ((j - i0) / LG + 1) * LG;
    if (i0 >= ntot) return;
    const double* lp = s.Lpub + (size_t)(j & 1) * kLpub;
    // every global operand of the first owned row block is requested before anything waits on one of them: the
    // published factor, the diagonal row block's previous column block, this block's previous column block, its
    // current column block, its look-ahead accumulator and (square part) its diagonal block -- ONE L2 round trip
    double rLi[2][2], rXB[2][2], rXA[2][2], rv[2][2], rw[2][2];
    {
        const double* xb = s.T + (size_t)j * NB * ld + (j > 0 ? (j - 1) * NB : 0);
        const double* xa = s.T + (size_t)i0 * NB * ld + (j > 0 ? (j - 1) * NB : 0);
        const double* o0 = s.T + (size_t)i0 * NB * ld + j * NB;
        const double* d0 = s.T + (size_t)i0 * NB * ld + (i0 < nsq ? i0 : 0) * NB;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int rr = ty + 16 * r, cc = tx + 16 * c;
                rLi[r][c] = ldg(lp + rr * NB + cc);
                rXB[r][c] = j > 0 ? ldg(xb + (size_t)rr * ld + cc) : 0.0;
                rXA[r][c] = j > 0 ? ldg(xa + (size_t)rr * ld + cc) : 0.0;
                rv[r][c] = ldg(o0 + (size_t)rr * ld + cc);
                if (j > 1) rv[r][c] -= ldg(s.Pacc + (size_t)i0 * NB * NB + rr * NB + cc);
                rw[r][c] = i0 < nsq ? ldg(d0 + (size_t)rr * ld + cc) : 0.0;
            }
        if (threadIdx.x < NB) sm.dinv[threadIdx.x] = ldg(lp + NB * NB + threadIdx.x);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                sm.Li[ty + 16 * r][tx + 16 * c] = rLi[r][c];
                sm.XB[ty + 16 * r][tx + 16 * c] = rXB[r][c];
                sm.XA[ty + 16 * r][tx + 16 * c] = rXA[r][c];
            }
    }
    for (int i = i0; i < ntot; i += LG) {
        double* out = s.T + (size_t)i * NB * ld + j * NB;
        double v[2][2];
        if (i == i0) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) v[r][c] = rv[r][c];
        } else {
            __syncthreads();
            if (j > 0) {
                const double* xa = s.T + (size_t)i * NB * ld + (j - 1) * NB;
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        sm.XA[ty + 16 * r][tx + 16 * c] = ldg(xa + (size_t)(ty + 16 * r) * ld + tx + 16 * c);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int rr = ty + 16 * r, cc = tx + 16 * c;
                    v[r][c] = ldg(out + (size_t)rr * ld + cc);
                    if (j > 1) v[r][c] -= ldg(s.Pacc + (size_t)i * NB * NB + rr * NB + cc);
                }
        }
        __syncthreads();
        if (j > 0) {
#pragma unroll 8
            for (int k = 0; k < NB; ++k) {
                const double a0 = sm.XA[ty][k], a1 = sm.XA[ty + 16][k], b0 = sm.XB[tx][k], b1 = sm.XB[tx + 16][k];
                v[0][0] = fma(-a0, b0, v[0][0]); v[0][1] = fma(-a0, b1, v[0][1]);
                v[1][0] = fma(-a1, b0, v[1][0]); v[1][1] = fma(-a1, b1, v[1][1]);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) sm.X[ty + 16 * r][tx + 16 * c] = v[r][c];
        __syncthreads();
        // X L_jj^-T by forward substitution, one lane per row: x_c = (v_c - sum_{q<c} x_q L[c][q]) / L[c][c]
        if (threadIdx.x < NB) {
            const int rr = threadIdx.x;
            double x[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double s0 = sm.X[rr][c], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int q = 0; q < c; ++q) {
                    const double l = sm.Li[c][q];
                    if ((q & 3) == 0) s0 = fma(-x[q], l, s0);
                    else if ((q & 3) == 1) s1 = fma(-x[q], l, s1);
                    else if ((q & 3) == 2) s2 = fma(-x[q], l, s2);
                    else s3 = fma(-x[q], l, s3);
                }
                x[c] = ((s0 + s1) + (s2 + s3)) * sm.dinv[c];
                sm.XO[rr][c] = x[c];
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) out[(size_t)(ty + 16 * r) * ld + tx + 16 * c] = sm.XO[ty + 16 * r][tx + 16 * c];
        if (i < nsq) {  // keep the own diagonal block current: T[i,i] -= X_ij X_ij^T
            double* dg = s.T + (size_t)i * NB * ld + i * NB;
            double w[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    w[r][c] = (i == i0) ? rw[r][c] : ldg(dg + (size_t)(ty + 16 * r) * ld + tx + 16 * c);
#pragma unroll 8
            for (int k = 0; k < NB; ++k) {
                const double a0 = sm.XO[ty][k], a1 = sm.XO[ty + 16][k], b0 = sm.XO[tx][k], b1 = sm.XO[tx + 16][k];
                w[0][0] = fma(-a0, b0, w[0][0]); w[0][1] = fma(-a0, b1, w[0][1]);
                w[1][0] = fma(-a1, b0, w[1][0]); w[1][1] = fma(-a1, b1, w[1][1]);
            }
            if (i == j + 1) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c) sm.L[ty + 16 * r][tx + 16 * c] = w[r][c];
                publish_factor(s, i, sm);
                if (i + LG < ntot) {  // more owned row blocks in this phase: bring back the inverse factor of block j
                    __syncthreads();
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            sm.Li[ty + 16 * r][tx + 16 * c] = ldg(lp + (ty + 16 * r) * NB + tx + 16 * c);
                    if (threadIdx.x < NB) sm.dinv[threadIdx.x] = ldg(lp + NB * NB + threadIdx.x);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c) dg[(size_t)(ty + 16 * r) * ld + tx + 16 * c] = w[r][c];
            }
        }
    }
    // look-ahead for phase j+1: Pacc[i] = sum_{p<j} X_ip X_{j+1,p}^T  (every operand final since phase j-1)
    if (j >= 1 && j + 1 < nsq) {
        long long tl = clock64();
        const bool last_owner = (ntot - 1) % LG == lb;
        for (int i = i0; i < ntot; i += LG) {
            if (i <= j + 1) continue;
            double acc[2][2];
            tile_nt<32, 32, false>(s.T + (size_t)i * NB * ld, ld, NB, s.T + (size_t)(j + 1) * NB * ld, ld, NB, j * NB, 0, 0,
                                   nullptr, acc, sm);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    s.Pacc[(size_t)i * NB * NB + trow<32>(r) * NB + tcol<32>(c)] = acc[r][c];
        }
        if (last_owner && threadIdx.x == 0) {
            const long long dtl = clock64() - tl;
            atomicAdd((unsigned long long*)&g_prof[6], (unsigned long long)dtl);
            atomicAdd((unsigned long long*)&g_prof[5], (unsigned long long)(-dtl));
        }
    }
}

// all panels of one or two independent systems of the same width (the second one -- the masked innovation
// covariance of the log-likelihood -- is swept in the same phases by the last quarter of the grid)
__device__ void chol_stack_grid(const ChSys& s1, const ChSys& s2, Smem& sm, GridSync& gs) {
    const int G = gridDim.x, b = blockIdx.x;
    const int G2 = s2.T ? (G / 4 > 0 ? G / 4 : 1) : 0, G1 = G - G2;
    const ChSys& s = (b < G1) ? s1 : s2;
    const int lb = (b < G1) ? b : b - G1, LG = (b < G1) ? G1 : G2;
    const int nsq = s1.n / NB;
    const int nb1 = s1.rows / NB, nb2 = s2.T ? s2.rows / NB : 0;
    const int npart = (G1 < nb1 ? G1 : nb1) + (G2 < nb2 ? G2 : nb2);   // CTAs that own row blocks
    const bool part = lb < ((b < G1) ? nb1 : nb2);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    if (part) {
        if (lb == 0) {  // diagonal block 0 needs no update
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    sm.L[ty + 16 * r][tx + 16 * c] = ldg(s.T + (size_t)(ty + 16 * r) * s.n + tx + 16 * c);
            publish_factor(s, 0, sm);
        }
        gs.sync_sub(npart);
        const bool last_owner = (b < G1) && ((s1.rows / NB - 1) % LG == lb);
        for (int j = 0; j < nsq; ++j) {
            long long t0 = clock64();
            chol_sys_phase(s, j, lb, LG, sm);
            if (last_owner) ST_PROF(5, t0);
            if (j + 1 < nsq) gs.sync_sub(npart);   // the full barrier below closes the last panel
            if (last_owner) ST_PROF(7, t0);
        }
    }
    gs.sync();
}

// ---- temporal block of the discretisation ----------------------------------------------------------------
template <int FAM>
struct TBlock {
    static constexpr int n = FamilyDim<FAM>::value;
    double A[n * n], Q[n * n], Pinf[n * n];
    __device__ void init(const bn_kernel_spec& sp, double h) {
        double Pp[symn(n)], X[n * n];
        MaternBlock<FAM, double>::transition(sp.lengthscale[0], h, A);
        MaternBlock<FAM, double>::pinf(sp.variance[0], sp.lengthscale[0], Pp);
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) Pinf[i * n + j] = Pp[sidx(i > j ? i : j, i > j ? j : i)];
        // Q = Pinf - A Pinf A^T (ops.py:149-151)
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) {
                double s = 0.0;
#pragma unroll
                for (int l = 0; l < n; ++l) s = fma(A[i * n + l], Pinf[l * n + j], s);
                X[i * n + j] = s;
            }
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) {
                double s = 0.0;
#pragma unroll
                for (int l = 0; l < n; ++l) s = fma(X[i * n + l], A[j * n + l], s);
                Q[i * n + j] = Pinf[i * n + j] - s;
            }
    }
    // out = A B A^T (+ Q on diagonal blocks)
    __device__ void rotate(const double* B, bool diag, double* out) const {
        double X[n * n];
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) {
                double s = 0.0;
#pragma unroll
                for (int l = 0; l < n; ++l) s = fma(A[i * n + l], B[l * n + j], s);
                X[i * n + j] = s;
            }
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) {
                double s = diag ? Q[i * n + j] : 0.0;
#pragma unroll
                for (int l = 0; l < n; ++l) s = fma(X[i * n + l], A[j * n + l], s);
                out[i * n + j] = s;
            }
    }
};

// ---- filter ----------------------------------------------------------------------------------------------
struct FilterArgs {
    bn_kernel_spec spec;
    int M;
    long long N;
    const double* dt;
    const double* y;      // [N,M]
    const double* R;      // [N,M,M]
    const uint8_t* mask;  // [N,M] nullable
    int return_predict;
    double* ell;          // nullable
    double* means;        // [N,d]
    double* covs;         // [N,d,d]
    double* T;            // [(Mp + dp + 32) x Mp]
    double* T2;           // [(Mp + 32) x Mp]   masked system of the log-likelihood (mask != null)
    double* Pacc;         // look-ahead accumulators of both systems
    double* Lpub;         // published inverse factors of both systems [2][2][32*32]
    double* Pcur;         // [d,d]   filtered covariance of the previous step when return_predict
    double* mcur;         // [d]
    double* ellacc;       // 1
    unsigned long long* ctr;
};

template <int FAM>
__global__ void __launch_bounds__(NTH) st_filter_kernel(FilterArgs a) {
    constexpr int n = FamilyDim<FAM>::value;
    extern __shared__ __align__(16) unsigned char smraw[];
    Smem& sm = *reinterpret_cast<Smem*>(smraw);
    GridSync gs{a.ctr, 0ULL};
    const int M = a.M, d = M * n, Mp = pad32(M), dp = pad32(d);
    const int rows = Mp + dp + NB;
    const long long tid = (long long)blockIdx.x * NTH + threadIdx.x, nthreads = (long long)gridDim.x * NTH;
    double* T = a.T;
    // padding of the stacked matrix: identity on the padded diagonal, zero elsewhere (the sweeps keep it so)
    for (long long e = tid; e < (long long)rows * Mp; e += nthreads) {
        const int r = (int)(e / Mp), c = (int)(e % Mp);
        T[e] = (r == c && r >= M) ? 1.0 : 0.0;
    }
    double* T2 = a.mask ? a.T2 : nullptr;
    const int rows2 = Mp + NB;
    if (T2)
        for (long long e = tid; e < (long long)rows2 * Mp; e += nthreads) {
            const int r = (int)(e / Mp), c = (int)(e % Mp);
            T2[e] = (r == c && r >= M) ? 1.0 : 0.0;
        }
    double ell_reg = 0.0;  // lives in lane 0 of warp 0 of the last CTA
    gs.sync();
    for (long long k = 0; k < a.N; ++k) {
        long long tp0 = clock64();
        // ---- predict + assemble: P^- = A P A^T + Q -> covs[k];  T = [H P^- H^T + R ; P^- H^T ; (y - H m^-)^T]
        TBlock<FAM> tb;
        tb.init(a.spec, a.dt[k]);
        const double* Pprev = a.return_predict ? a.Pcur : a.covs + (size_t)(k > 0 ? k - 1 : 0) * d * d;
        const double* mprev = a.return_predict ? a.mcur : a.means + (size_t)(k > 0 ? k - 1 : 0) * d;
        double* Pk = a.covs + (size_t)k * d * d;
        double* mk = a.means + (size_t)k * d;
        const double* Rk = a.R + (size_t)k * M * M;
        for (long long e = tid; e < (long long)M * M; e += nthreads) {
            const int i = (int)(e / M), j = (int)(e % M);
            double B[n * n], O[n * n];
#pragma unroll
            for (int p = 0; p < n; ++p)
#pragma unroll
                for (int q = 0; q < n; ++q)
                    B[p * n + q] = (k == 0) ? (i == j ? tb.Pinf[p * n + q] : 0.0) : ldg(Pprev + (size_t)(i * n + p) * d + j * n + q);
            tb.rotate(B, i == j, O);
#pragma unroll
            for (int p = 0; p < n; ++p)
#pragma unroll
                for (int q = 0; q < n; ++q) Pk[(size_t)(i * n + p) * d + j * n + q] = O[p * n + q];
            const double Sij = O[0] + Rk[(size_t)i * M + j];
            T[(size_t)i * Mp + j] = Sij;
            if (T2) {  // mvn_logpdf with a mask (utils.py:382-388): masked rows/columns independent, variance 1/(2 pi)
                const bool mi = a.mask[(size_t)k * M + i] != 0, mj = a.mask[(size_t)k * M + j] != 0;
                T2[(size_t)i * Mp + j] = (mi || mj) ? ((i == j) ? 0.15915494309189535 : 0.0) : Sij;
            }
#pragma unroll
            for (int p = 0; p < n; ++p) T[(size_t)(Mp + i * n + p) * Mp + j] = O[p * n];
        }
        for (long long i = tid; i < M; i += nthreads) {
            double mp[n];
#pragma unroll
            for (int p = 0; p < n; ++p) {
                double s = 0.0;
                if (k > 0) {
#pragma unroll
                    for (int q = 0; q < n; ++q) s = fma(tb.A[p * n + q], ldg(mprev + i * n + q), s);
                }
                mp[p] = s;
                mk[i * n + p] = s;
            }
            const double res = a.y[(size_t)k * M + i] - mp[0];
            T[(size_t)(Mp + dp) * Mp + i] = res;
            if (T2) T2[(size_t)Mp * Mp + i] = a.mask[(size_t)k * M + i] ? 0.0 : res;
        }
        if (blockIdx.x == 0) ST_PROF(0, tp0);
        gs.sync();
        if (blockIdx.x == 0) ST_PROF(1, tp0);
        // ---- S = L L^T;  X = P^- H^T L^-T;  z = L^-1 (y - H m^-)
        {
            const ChSys s1{T, Mp, rows, a.Pacc, a.Lpub};
            const ChSys s2{T2, Mp, rows2, a.Pacc + (size_t)(rows / NB) * NB * NB, a.Lpub + 2 * kLpub};
            chol_stack_grid(s1, s2, sm, gs);
        }
        if (blockIdx.x == 0) ST_PROF(2, tp0);
        // ---- P = P^- - X X^T;  m = m^- + X z;  ell += log N(y | H m^-, S)   (ops.py:163-172)
        const double* X = T + (size_t)Mp * Mp;
        const double* z = T + (size_t)(Mp + dp) * Mp;
        double* Pout = a.return_predict ? a.Pcur : Pk;
        double* mout = a.return_predict ? a.mcur : mk;
        const int tm = (d + 31) / 32, tn = (d + 63) / 64;
        for (int t = blockIdx.x; t < tm * tn; t += gridDim.x) {
            const int i0 = (t / tn) * 32, j0 = (t % tn) * 64;
            double acc[2][4];
            tile_nt<32, 64, false>(X, Mp, d, X, Mp, d, Mp, i0, j0, nullptr, acc, sm);
            const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int row = i0 + trow<32>(r), col = j0 + tcol<64>(c);
                    if (row < d && col < d) Pout[(size_t)row * d + col] = ldg(Pk + (size_t)row * d + col) - acc[r][c];
                }
        }
        {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            for (int row = blockIdx.x * (NTH / 32) + warp; row < d; row += gridDim.x * (NTH / 32)) {
                double s = 0.0;
                for (int c = lane; c < M; c += 32) s = fma(ldg(X + (size_t)row * Mp + c), ldg(z + c), s);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (lane == 0) mout[row] = ldg(mk + row) + s;
            }
            if (blockIdx.x == gridDim.x - 1 && warp == 0) {
                double q = 0.0, ld = 0.0;
                const double* Tl = T2 ? T2 : T;
                const double* zl = T2 ? T2 + (size_t)Mp * Mp : z;
                for (int c = lane; c < M; c += 32) {
                    const double zc = ldg(zl + c);
                    q = fma(zc, zc, q);
                    ld += log(ldg(Tl + (size_t)c * Mp + c));
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    q += __shfl_xor_sync(0xffffffffu, q, off);
                    ld += __shfl_xor_sync(0xffffffffu, ld, off);
                }
                if (lane == 0) ell_reg += -0.5 * (q + M * 1.8378770664093453 + 2.0 * ld);
            }
        }
        if (blockIdx.x == 0) ST_PROF(3, tp0);
        gs.sync();
        if (blockIdx.x == 0) ST_PROF(4, tp0);
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && a.ell) *a.ell = ell_reg;
}

// ---- mean-field filter / smoother ------------------------------------------------------
// kalman_filter_meanfield / rauch_tung_striebel_smoother_meanfield (ops.py:429-467, 581-611, 614-650, 681-706): the
// state covariance is truncated to its M diagonal n x n blocks after every update, so the state is M blocks
// (means [N,M,n,1], covs [N,M,n,n]) and a step needs of the M x M innovation covariance S = diag(P^-_i[0,0]) + R only
//     S^-1 r,   diag(S^-1),   log det S:
// P_i = P^-_i - c_i (S^-1)_ii c_i^T,  m_i = m^-_i + c_i (S^-1 r)_i  with c_i = P^-_i[:,0].  The stacked sweep of
// T = [S ; I ; r^T] gives LiT = L^-T and z = L^-1 r, hence (S^-1 r)_i = LiT[i,:] z and (S^-1)_ii = |LiT[i,:]|^2.
struct MfFilterArgs {
    bn_kernel_spec spec;
    int M;
    long long N;
    const double* dt;
    const double* y;      // [N,M]
    const double* R;      // [N,M,M]
    const uint8_t* mask;  // [N,M] nullable
    double* ell;          // nullable
    double* means;        // [N,M,n]
    double* covs;         // [N,M,n,n]
    double* T;            // [(2 Mp + 32) x Mp]
    double* T2;           // [(Mp + 32) x Mp]
    double* Pacc;
    double* Lpub;
    unsigned long long* ctr;
};

template <int FAM>
__global__ void __launch_bounds__(NTH) st_mf_filter_kernel(MfFilterArgs a) {
    constexpr int n = FamilyDim<FAM>::value;
    extern __shared__ __align__(16) unsigned char smraw[];
    Smem& sm = *reinterpret_cast<Smem*>(smraw);
    GridSync gs{a.ctr, 0ULL};
    const int M = a.M, Mp = pad32(M);
    const int rows = 2 * Mp + NB, rows2 = Mp + NB;
    const long long tid = (long long)blockIdx.x * NTH + threadIdx.x, nthreads = (long long)gridDim.x * NTH;
    double* T = a.T;
    double* T2 = a.mask ? a.T2 : nullptr;
    for (long long e = tid; e < (long long)rows * Mp; e += nthreads) {
        const int r = (int)(e / Mp), c = (int)(e % Mp);
        T[e] = ((r == c && r >= M && r < Mp) || (r - Mp == c && r >= Mp && r < 2 * Mp)) ? 1.0 : 0.0;
    }
    if (T2)
        for (long long e = tid; e < (long long)rows2 * Mp; e += nthreads) {
            const int r = (int)(e / Mp), c = (int)(e % Mp);
            T2[e] = (r == c && r >= M) ? 1.0 : 0.0;
        }
    double ell_reg = 0.0;
    gs.sync();
    for (long long k = 0; k < a.N; ++k) {
        TBlock<FAM> tb;
        tb.init(a.spec, a.dt[k]);
        double* Pk = a.covs + (size_t)k * M * n * n;
        double* mk = a.means + (size_t)k * M * n;
        const double* Pprev = Pk - (size_t)M * n * n;
        const double* mprev = mk - (size_t)M * n;
        const double* Rk = a.R + (size_t)k * M * M;
        // ---- predict every block, assemble T = [diag(P^-_i[0,0]) + R ; I ; (y - H m^-)^T]
        for (long long e = tid; e < (long long)M * M; e += nthreads) {
            const int i = (int)(e / M), j = (int)(e % M);
            double Sij = Rk[(size_t)i * M + j];
            double res = 0.0;
            if (i == j) {
                double B[n * n], O[n * n], mp[n];
#pragma unroll
                for (int p = 0; p < n * n; ++p) B[p] = (k == 0) ? tb.Pinf[p] : ldg(Pprev + (size_t)i * n * n + p);
                tb.rotate(B, true, O);
#pragma unroll
                for (int p = 0; p < n * n; ++p) Pk[(size_t)i * n * n + p] = O[p];
#pragma unroll
                for (int p = 0; p < n; ++p) {
                    double sv = 0.0;
                    if (k > 0) {
#pragma unroll
                        for (int q = 0; q < n; ++q) sv = fma(tb.A[p * n + q], ldg(mprev + i * n + q), sv);
                    }
                    mp[p] = sv;
                    mk[i * n + p] = sv;
                }
                Sij += O[0];
                res = a.y[(size_t)k * M + i] - mp[0];
                T[(size_t)(2 * Mp) * Mp + i] = res;
                if (T2) T2[(size_t)Mp * Mp + i] = a.mask[(size_t)k * M + i] ? 0.0 : res;
            }
            T[(size_t)i * Mp + j] = Sij;
            T[(size_t)(Mp + i) * Mp + j] = (i == j) ? 1.0 : 0.0;
            if (T2) {
                const bool mi = a.mask[(size_t)k * M + i] != 0, mj = a.mask[(size_t)k * M + j] != 0;
                T2[(size_t)i * Mp + j] = (mi || mj) ? ((i == j) ? 0.15915494309189535 : 0.0) : Sij;
            }
        }
        gs.sync();
        {
            const ChSys s1{T, Mp, rows, a.Pacc, a.Lpub};
            const ChSys s2{T2, Mp, rows2, a.Pacc + (size_t)(rows / NB) * NB * NB, a.Lpub + 2 * kLpub};
            chol_stack_grid(s1, s2, sm, gs);
        }
        // ---- per block: (S^-1)_ii and (S^-1 r)_i from row i of LiT, then the rank-one update of the block
        const double* LiT = T + (size_t)Mp * Mp;
        const double* z = T + (size_t)(2 * Mp) * Mp;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int i = blockIdx.x * (NTH / 32) + warp; i < M; i += gridDim.x * (NTH / 32)) {
            double sii = 0.0, gi = 0.0;
            for (int c = lane; c < M; c += 32) {
                const double l = ldg(LiT + (size_t)i * Mp + c);
                sii = fma(l, l, sii);
                gi = fma(l, ldg(z + c), gi);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                sii += __shfl_xor_sync(0xffffffffu, sii, off);
                gi += __shfl_xor_sync(0xffffffffu, gi, off);
            }
            if (lane == 0) {
                double P[n * n], c0[n];
#pragma unroll
                for (int p = 0; p < n * n; ++p) P[p] = ldg(Pk + (size_t)i * n * n + p);
#pragma unroll
                for (int p = 0; p < n; ++p) c0[p] = P[p * n];
#pragma unroll
                for (int p = 0; p < n; ++p) {
                    mk[i * n + p] = ldg(mk + i * n + p) + c0[p] * gi;
#pragma unroll
                    for (int q = 0; q < n; ++q) Pk[(size_t)i * n * n + p * n + q] = P[p * n + q] - c0[p] * sii * P[q];
                }
            }
        }
        if (blockIdx.x == gridDim.x - 1 && warp == 0) {
            double q = 0.0, ld = 0.0;
            const double* Tl = T2 ? T2 : T;
            const double* zl = T2 ? T2 + (size_t)Mp * Mp : z;
            for (int c = lane; c < M; c += 32) {
                const double zc = ldg(zl + c);
                q = fma(zc, zc, q);
                ld += log(ldg(Tl + (size_t)c * Mp + c));
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                q += __shfl_xor_sync(0xffffffffu, q, off);
                ld += __shfl_xor_sync(0xffffffffu, ld, off);
            }
            if (lane == 0) ell_reg += -0.5 * (q + M * 1.8378770664093453 + 2.0 * ld);
        }
        gs.sync();
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && a.ell) *a.ell = ell_reg;
}

// independent RTS pass per block: one thread per block, time sequential (ops.py:614-650)
template <int FAM>
__global__ void st_mf_smoother_kernel(bn_kernel_spec spec, int M, long long N, const double* dt, const double* fm,
                                      const double* fP, int return_full, double* means, double* covs, double* gains) {
    constexpr int n = FamilyDim<FAM>::value;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double sm[n], sP[n * n];
#pragma unroll
    for (int p = 0; p < n; ++p) sm[p] = fm[((size_t)(N - 1) * M + i) * n + p];
#pragma unroll
    for (int p = 0; p < n * n; ++p) sP[p] = fP[((size_t)(N - 1) * M + i) * n * n + p];
    for (long long k = N - 1; k >= 0; --k) {
        TBlock<FAM> tb;
        tb.init(spec, dt[k]);
        double f[n], F[n * n], pm[n], AfP[n * n], pP[n * n], Lp[symn(n)], Ct[n * n];
#pragma unroll
        for (int p = 0; p < n; ++p) f[p] = fm[((size_t)k * M + i) * n + p];
#pragma unroll
        for (int p = 0; p < n * n; ++p) F[p] = fP[((size_t)k * M + i) * n * n + p];
        matvec<n, n>(tb.A, f, pm);
        matmul<n, n, n>(tb.A, F, AfP);
#pragma unroll
        for (int p = 0; p < n; ++p)
#pragma unroll
            for (int q = 0; q < n; ++q) {
                double sv = tb.Q[p * n + q];
#pragma unroll
                for (int l = 0; l < n; ++l) sv = fma(AfP[p * n + l], tb.A[q * n + l], sv);
                pP[p * n + q] = sv;
            }
#pragma unroll
        for (int p = 0; p < n; ++p)
#pragma unroll
            for (int q = 0; q <= p; ++q) Lp[sidx(p, q)] = pP[p * n + q];
        chol<n>(Lp);
#pragma unroll
        for (int p = 0; p < n * n; ++p) Ct[p] = AfP[p];   // C^T = pP^-1 (A fP)
        chol_solve<n, n>(Lp, Ct);
        double dm[n], D[n * n], CD[n * n], nm[n], nP[n * n];
#pragma unroll
        for (int p = 0; p < n; ++p) dm[p] = sm[p] - pm[p];
#pragma unroll
        for (int p = 0; p < n * n; ++p) D[p] = sP[p] - pP[p];
#pragma unroll
        for (int p = 0; p < n; ++p) {
            double sv = f[p];
#pragma unroll
            for (int l = 0; l < n; ++l) sv = fma(Ct[l * n + p], dm[l], sv);   // C[p][l] = Ct[l][p]
            nm[p] = sv;
#pragma unroll
            for (int q = 0; q < n; ++q) {
                double t = 0.0;
#pragma unroll
                for (int l = 0; l < n; ++l) t = fma(Ct[l * n + p], D[l * n + q], t);
                CD[p * n + q] = t;
            }
        }
#pragma unroll
        for (int p = 0; p < n; ++p)
#pragma unroll
            for (int q = 0; q < n; ++q) {
                double sv = F[p * n + q];
#pragma unroll
                for (int l = 0; l < n; ++l) sv = fma(CD[p * n + l], Ct[l * n + q], sv);
                nP[p * n + q] = sv;
            }
#pragma unroll
        for (int p = 0; p < n; ++p) sm[p] = nm[p];
#pragma unroll
        for (int p = 0; p < n * n; ++p) sP[p] = nP[p];
        if (gains) {
#pragma unroll
            for (int p = 0; p < n; ++p)
#pragma unroll
                for (int q = 0; q < n; ++q) gains[((size_t)k * M + i) * n * n + p * n + q] = Ct[q * n + p];
        }
        if (return_full) {
#pragma unroll
            for (int p = 0; p < n; ++p) means[((size_t)k * M + i) * n + p] = sm[p];
#pragma unroll
            for (int p = 0; p < n * n; ++p) covs[((size_t)k * M + i) * n * n + p] = sP[p];
        } else {
            means[(size_t)k * M + i] = sm[0];
            covs[(size_t)k * M * M + (size_t)i * M + i] = sP[0];
        }
    }
}

// ---- smoother ----------------------------------------------------------------------------------------------
// The smoother gain G_k = fP_k A^T (A fP_k A^T + Q)^-1 (ops.py:293-296) depends on the FILTER output only, not on
// the backward recursion, so the factorisations of all time steps are independent: a batched kernel (one CTA per
// time step) forms every gain, and the sequential part that remains is two tile products per step,
//     sP_k = fP_k + G_k (sP_{k+1} - P^-_k) G_k^T,    sm_k = fm_k + G_k (sm_{k+1} - A fm_k)      (ops.py:297-298).
struct GainArgs {
    bn_kernel_spec spec;
    int M;
    long long k0, k1;     // time steps [k0, k1)
    const double* dt;     // step OUT OF k (basemodels.py:700)
    const double* fP;     // [N,d,d]
    double* G;            // gains of step k at G + (k - k0) d d
    double* T;            // per-CTA slots [3 dp x dp]
};

template <int FAM>
__global__ void __launch_bounds__(NTH) st_gain_kernel(GainArgs a) {
    constexpr int n = FamilyDim<FAM>::value;
    extern __shared__ __align__(16) unsigned char smraw[];
    Smem& sm = *reinterpret_cast<Smem*>(smraw);
    const int M = a.M, d = M * n, dp = pad32(d), rows = 3 * dp;
    double* T = a.T + (size_t)blockIdx.x * rows * dp;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int e = threadIdx.x; e < rows * dp; e += NTH) {
        const int r = e / dp, c = e % dp;
        T[e] = (r == c && r >= d && r < dp) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (long long k = a.k0 + blockIdx.x; k < a.k1; k += gridDim.x) {
        TBlock<FAM> tb;
        tb.init(a.spec, a.dt[k]);
        const double* fPk = a.fP + (size_t)k * d * d;
        // T = [P^- ; fP A^T ; I]
        for (int e = threadIdx.x; e < M * M; e += NTH) {
            const int i = e / M, j = e % M;
            double B[n * n], O[n * n];
#pragma unroll
            for (int p = 0; p < n; ++p)
#pragma unroll
                for (int q = 0; q < n; ++q) B[p * n + q] = fPk[(size_t)(i * n + p) * d + j * n + q];
            tb.rotate(B, i == j, O);
#pragma unroll
            for (int p = 0; p < n; ++p)
#pragma unroll
                for (int q = 0; q < n; ++q) {
                    const size_t r = i * n + p, c = j * n + q;
                    T[r * dp + c] = O[p * n + q];
                    double sv = 0.0;  // (fP A^T)[r][c] = sum_l fP[r][j n + l] A_t[q][l]
#pragma unroll
                    for (int l = 0; l < n; ++l) sv = fma(B[p * n + l], tb.A[q * n + l], sv);
                    T[(size_t)(dp + r) * dp + c] = sv;
                    T[(size_t)(2 * dp + r) * dp + c] = (r == c) ? 1.0 : 0.0;
                }
        }
        __syncthreads();
        __threadfence_block();
        // P^- = L L^T;  Y = fP A^T L^-T;  LiT = L^-T
        chol_stack_cta(T, dp, rows, sm, 2 * dp / NB);
        // G = Y L^-1 = Y LiT^T-as-rows  (NT form)
        const double* Y = T + (size_t)dp * dp;
        const double* LiT = T + (size_t)2 * dp * dp;
        double* G = a.G + (size_t)(k - a.k0) * d * d;
        const int tm = (d + 63) / 64;
        for (int t = 0; t < tm * tm; ++t) {
            const int i0 = (t / tm) * 64, j0 = (t % tm) * 64;
            double acc[4][4];
            // LiT row j is zero left of column j: start the k-loop at the tile's first row (multiple of 32)
            tile_nt<64, 64, false>(Y + j0, dp, d, LiT + j0, dp, d, dp - j0, i0, j0, nullptr, acc, sm);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int row = i0 + trow<64>(r), col = j0 + tcol<64>(c);
                    if (row < d && col < d) G[(size_t)row * d + col] = acc[r][c];
                }
        }
        __syncthreads();
    }
}

struct SmootherArgs {
    bn_kernel_spec spec;
    int M;
    long long N;
    long long k0, k1;     // this launch runs k = k1-1 .. k0
    int first;            // 1: start the recursion from the last filtered state (ops.py:304-305)
    const double* dt;
    const double* fm;     // [N,d]
    const double* fP;     // [N,d,d]
    int return_full;
    double* means;        // [N,M] or [N,d]
    double* covs;         // [N,M,M] or [N,d,d]
    const double* G;      // gains of step k at G + (k - k0) d d
    double* sP;           // [d,d]  full smoothed covariance of the step done last (carried between launches)
    double* smv;          // [d]
    double* Dm;           // [d,d]  sP_{k+1} - P^-_k
    double* v;            // [d]    sm_{k+1} - A_k fm_k
    double* Z;            // [d,d]
    unsigned long long* ctr;
};

// P^-_k[row][col] and (A_k fm_k)[row] from the filtered state of step k, element-wise
template <int FAM>
__device__ __forceinline__ double ppred_elem(const TBlock<FAM>& tb, const double* fPk, int d, int row, int col) {
    constexpr int n = FamilyDim<FAM>::value;
    const int i = row / n, pa = row % n, j = col / n, pb = col % n;
    double sv = (i == j) ? tb.Q[pa * n + pb] : 0.0;
#pragma unroll
    for (int p = 0; p < n; ++p) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < n; ++q) t = fma(fPk[(size_t)(i * n + p) * d + j * n + q], tb.A[pb * n + q], t);
        sv = fma(tb.A[pa * n + p], t, sv);
    }
    return sv;
}
template <int FAM>
__device__ __forceinline__ double mpred_elem(const TBlock<FAM>& tb, const double* fmk, int row) {
    constexpr int n = FamilyDim<FAM>::value;
    const int i = row / n, pa = row % n;
    double sv = 0.0;
#pragma unroll
    for (int q = 0; q < n; ++q) sv = fma(tb.A[pa * n + q], fmk[i * n + q], sv);
    return sv;
}

template <int FAM>
__global__ void __launch_bounds__(NTH) st_smoother_kernel(SmootherArgs a) {
    constexpr int n = FamilyDim<FAM>::value;
    extern __shared__ __align__(16) unsigned char smraw[];
    Smem& sm = *reinterpret_cast<Smem*>(smraw);
    GridSync gs{a.ctr, 0ULL};
    const int M = a.M, d = M * n;
    const long long tid = (long long)blockIdx.x * NTH + threadIdx.x, nthreads = (long long)gridDim.x * NTH;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tm = (d + 31) / 32, tn = (d + 63) / 64;
    {   // Dm, v of the first step of this launch from the carried (or last filtered) state
        const long long k = a.k1 - 1;
        TBlock<FAM> tb;
        tb.init(a.spec, a.dt[k]);
        const double* fPk = a.fP + (size_t)k * d * d;
        const double* fmk = a.fm + (size_t)k * d;
        const double* sPn = a.first ? a.fP + (size_t)(a.N - 1) * d * d : a.sP;
        const double* smn = a.first ? a.fm + (size_t)(a.N - 1) * d : a.smv;
        for (long long e = tid; e < (long long)d * d; e += nthreads) {
            const int row = (int)(e / d), col = (int)(e % d);
            a.Dm[e] = sPn[e] - ppred_elem<FAM>(tb, fPk, d, row, col);
        }
        for (long long e = tid; e < d; e += nthreads) a.v[e] = smn[e] - mpred_elem<FAM>(tb, fmk, (int)e);
    }
    gs.sync();
    for (long long k = a.k1 - 1; k >= a.k0; --k) {
        long long tp0 = clock64();
        const double* fPk = a.fP + (size_t)k * d * d;
        const double* fmk = a.fm + (size_t)k * d;
        const double* G = a.G + (size_t)(k - a.k0) * d * d;
        const bool more = k > 0;  // a step k-1 exists (possibly in the next launch): prepare its Dm, v
        TBlock<FAM> tb;
        if (more) tb.init(a.spec, a.dt[k - 1]);
        // ---- Z = G (sP_next - P^-)  (Dm symmetric: NT form);  sm = fm + G v
        for (int t = blockIdx.x; t < tm * tn; t += gridDim.x) {
            const int i0 = (t / tn) * 32, j0 = (t % tn) * 64;
            double acc[2][4];
            tile_nt<32, 64, false>(G, d, d, a.Dm, d, d, d, i0, j0, nullptr, acc, sm);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int row = i0 + trow<32>(r), col = j0 + tcol<64>(c);
                    if (row < d && col < d) a.Z[(size_t)row * d + col] = acc[r][c];
                }
        }
        for (int row = blockIdx.x * (NTH / 32) + warp; row < d; row += gridDim.x * (NTH / 32)) {
            double sv = 0.0;
            for (int c = lane; c < d; c += 32) sv = fma(G[(size_t)row * d + c], ldg(a.v + c), sv);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, off);
            if (lane == 0) {
                const double val = fmk[row] + sv;
                a.smv[row] = val;
                if (a.return_full) a.means[(size_t)k * d + row] = val;
                else if (row % n == 0) a.means[(size_t)k * M + row / n] = val;
            }
        }
        if (blockIdx.x == 0) ST_PROF(3, tp0);
        gs.sync();
        if (blockIdx.x == 0) ST_PROF(4, tp0);
        // ---- sP = fP + Z G^T (ops.py:298);  next step's Dm = sP - P^-_{k-1},  v = sm - A_{k-1} fm_{k-1}
        const double* fPm = fPk - (size_t)d * d;
        for (int t = blockIdx.x; t < tm * tn; t += gridDim.x) {
            const int i0 = (t / tn) * 32, j0 = (t % tn) * 64;
            double acc[2][4];
            tile_nt<32, 64, false>(a.Z, d, d, G, d, d, d, i0, j0, nullptr, acc, sm);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int row = i0 + trow<32>(r), col = j0 + tcol<64>(c);
                    if (row < d && col < d) {
                        const double val = fPk[(size_t)row * d + col] + acc[r][c];
                        if (k == a.k0) a.sP[(size_t)row * d + col] = val;  // carried to the next launch
                        if (more) a.Dm[(size_t)row * d + col] = val - ppred_elem<FAM>(tb, fPm, d, row, col);
                        if (a.return_full) a.covs[(size_t)k * d * d + (size_t)row * d + col] = val;
                        else if (row % n == 0 && col % n == 0)
                            a.covs[(size_t)k * M * M + (size_t)(row / n) * M + col / n] = val;
                    }
                }
        }
        if (more)
            for (long long e = tid; e < d; e += nthreads) a.v[e] = ldg(a.smv + e) - mpred_elem<FAM>(tb, fmk - d, (int)e);
        if (blockIdx.x == 0) ST_PROF(3, tp0);
        gs.sync();
        if (blockIdx.x == 0) ST_PROF(4, tp0);
    }
}

// ---- batched SPD inverse / pseudo-likelihood projection ------------------------------------------------------
// One CTA per time step.  S_k = (PROJECT ? Bt diag(lam_k) Bt^T : A_k) + jitter I;  out_k = S_k^-1 through the
// Cholesky factor (utils.py:22-35);  rhs -> S_k^-1 rhs;  logdet -> log det S_k.
struct InvArgs {
    long long N;
    int n;                // matrix size (M)
    int Ns;               // PROJECT: inner dimension
    const double* A;      // [N,n,n]        (!PROJECT)
    const double* Bt;     // [n,Ns]         (PROJECT)  B^T, time-invariant
    const double* lam;    // [N,Ns]         (PROJECT)  diagonal of nat2
    const double* rhs_s;  // [N,Ns]         (PROJECT)  nat1 (projected by Bt first) | [N,n] (!PROJECT) | null
    double jitter;
    double* S;            // [N,n,n] nullable: the matrix that was inverted (nat2_full)
    double* inv;          // [N,n,n]
    double* sol;          // [N,n] nullable
    double* logdet;       // [N] nullable
    double* T;            // per-CTA slots [(2 np + 32) x np]
};

template <bool PROJECT>
__global__ void __launch_bounds__(NTH) st_inverse_kernel(InvArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    Smem& sm = *reinterpret_cast<Smem*>(smraw);
    const int n = a.n, np = pad32(n), rows = 2 * np + NB;
    double* T = a.T + (size_t)blockIdx.x * rows * np;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int tm = (n + 63) / 64;
    for (long long k = blockIdx.x; k < a.N; k += gridDim.x) {
        for (int e = threadIdx.x; e < rows * np; e += NTH) {
            const int r = e / np, c = e % np;
            double v = 0.0;
            if (r < np) {
                if (r == c) v = (r < n) ? a.jitter : 1.0;
                if (!PROJECT && r < n && c < n) v += a.A[(size_t)k * n * n + (size_t)r * n + c];
            } else if (r < 2 * np) {
                v = (r - np == c) ? 1.0 : 0.0;
            } else if (r == 2 * np && c < n && a.rhs_s) {
                if (!PROJECT) v = a.rhs_s[(size_t)k * n + c];
            }
            T[e] = v;
        }
        __syncthreads();
        if (PROJECT) {
            // S = Bt diag(lam) Bt^T (basemodels.py:681), rhs = Bt nat1 (:680)
            const double* lam = a.lam + (size_t)k * a.Ns;
            for (int t = 0; t < tm * tm; ++t) {
                const int i0 = (t / tm) * 64, j0 = (t % tm) * 64;
                if (j0 > i0) continue;  // lower triangle + mirror
                double acc[4][4];
                tile_nt<64, 64, true>(a.Bt, a.Ns, n, a.Bt, a.Ns, n, a.Ns, i0, j0, lam, acc, sm);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int row = i0 + trow<64>(r), col = j0 + tcol<64>(c);
                        if (row < n && col < n) {
                            T[(size_t)row * np + col] += acc[r][c];
                            if (j0 < i0) T[(size_t)col * np + row] += acc[r][c];
                        }
                    }
            }
            if (a.rhs_s) {
                const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
                const double* r1 = a.rhs_s + (size_t)k * a.Ns;
                for (int row = warp; row < n; row += NTH / 32) {
                    double s = 0.0;
                    for (int c = lane; c < a.Ns; c += 32) s = fma(a.Bt[(size_t)row * a.Ns + c], r1[c], s);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                    if (lane == 0) T[(size_t)(2 * np) * np + row] = s;
                }
            }
            __syncthreads();
            __threadfence_block();
        }
        if (a.S) {
            for (int e = threadIdx.x; e < n * n; e += NTH) {
                const int r = e / n, c = e % n;
                a.S[(size_t)k * n * n + e] = ldg(T + (size_t)r * np + c) - (r == c ? a.jitter : 0.0);
            }
        }
        __syncthreads();
        chol_stack_cta(T, np, rows, sm, np / NB);
        // inverse = L^-T L^-1 = LiT LiT^T with LiT = T[np:2np]  (cho_solve(L, I))
        const double* LiT = T + (size_t)np * np;
        for (int t = 0; t < tm * tm; ++t) {
            const int i0 = (t / tm) * 64, j0 = (t % tm) * 64;
            if (j0 > i0) continue;
            double acc[4][4];
            tile_nt<64, 64, false>(LiT, np, n, LiT, np, n, np, i0, j0, nullptr, acc, sm);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int row = i0 + trow<64>(r), col = j0 + tcol<64>(c);
                    if (row < n && col < n) {
                        a.inv[(size_t)k * n * n + (size_t)row * n + col] = acc[r][c];
                        if (j0 < i0) a.inv[(size_t)k * n * n + (size_t)col * n + row] = acc[r][c];
                    }
                }
        }
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (a.sol) {
            const double* z = T + (size_t)(2 * np) * np;
            for (int row = warp; row < n; row += NTH / 32) {
                double s = 0.0;
                for (int c = lane; c < n; c += 32) s = fma(ldg(LiT + (size_t)row * np + c), ldg(z + c), s);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (lane == 0) a.sol[(size_t)k * n + row] = s;
            }
        }
        if (a.logdet && warp == 0) {
            double ld = 0.0;
            for (int c = lane; c < n; c += 32) ld += log(ldg(T + (size_t)c * np + c));
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, off);
            if (lane == 0) a.logdet[k] = 2.0 * ld;
        }
        __syncthreads();
    }
}

// ---- conditional_posterior_to_data (basemodels.py:743-764), marginals only --------------------------------------
// mean_f = B m,  var_f = diag(B V B^T) + cdiag.  One CTA per time step; B [Ns,M] time-invariant.
__global__ void __launch_bounds__(NTH) st_to_data_kernel(long long N, int Ns, int M, const double* B, const double* cdiag,
                                                         const double* pm, const double* pV, double* mean_f,
                                                         double* var_f) {
    extern __shared__ __align__(16) unsigned char smraw[];
    Smem& sm = *reinterpret_cast<Smem*>(smraw);
    double* part = &sm.L[0][0];  // 64 x 16 partial sums (fits the 32 x 33 block buffer)
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (long long k = blockIdx.x; k < N; k += gridDim.x) {
        const double* V = pV + (size_t)k * M * M;
        const double* m = pm + (size_t)k * M;
        for (int i0 = 0; i0 < Ns; i0 += 64) {
            double dsum[4] = {0.0, 0.0, 0.0, 0.0};
            for (int j0 = 0; j0 < M; j0 += 64) {
                double acc[4][4];  // (B V)[i0.., j0..]  (V symmetric: NT form)
                tile_nt<64, 64, false>(B, M, Ns, V, M, M, M, i0, j0, nullptr, acc, sm);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int row = i0 + trow<64>(r), col = j0 + tcol<64>(c);
                        if (row < Ns && col < M) dsum[r] = fma(acc[r][c], B[(size_t)row * M + col], dsum[r]);
                    }
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 4; ++r) part[trow<64>(r) * 16 + ((threadIdx.x >> 5) & 3) * 4 + (threadIdx.x & 3)] = dsum[r];
            __syncthreads();
            if (threadIdx.x < 64) {
                const int row = i0 + threadIdx.x;
                if (row < Ns) {
                    double s = 0.0;
#pragma unroll
                    for (int q = 0; q < 16; ++q) s += part[threadIdx.x * 16 + q];
                    var_f[(size_t)k * Ns + row] = s + (cdiag ? cdiag[row] : 0.0);
                    double mu = 0.0;
                    for (int c = 0; c < M; ++c) mu = fma(B[(size_t)row * M + c], m[c], mu);
                    mean_f[(size_t)k * Ns + row] = mu;
                }
            }
            __syncthreads();
        }
    }
}

// ---- prediction at test times (MarkovGaussianProcess.predict, basemodels.py:766-816, for a SpatioTemporalKernel) -------
// temporal_conditional (utils.py:99-136, 173-215) on the Kronecker state: with the n x n temporal conditional
// [P1, W], T of the two gaps around the test time, the state covariance is
//     (I (x) P1) C_- (I (x) P1)^T + (I (x) P1) X (I (x) W)^T + (I (x) W) X^T (I (x) P1)^T + (I (x) W) C_+ (I (x) W)^T + I (x) T,
// X = G C_+ (gain of the left neighbour times the covariance of the right one: the only dense product), and what
// predict() needs is H (.) H^T, i.e. per block pair the quadratic forms with p = P1[0,:], w = W[0,:].  One CTA per
// test time; C_+ = Pinf-Kronecker (pk) and G = 0 at the ends (the dummy states of basemodels.py:793-794).
struct StPredArgs {
    bn_kernel_spec spec;
    int M;
    long long N, Nq;
    const double* x;      // [N] training times
    const double* xs;     // [Nq] test times
    const double* sm;     // [N,d]
    const double* sP;     // [N,d,d]
    const double* gain;   // [N,d,d]
    const double* pk;     // [d,d]  I (x) Pinf_t
    double* fmean;        // [Nq,M]
    double* fcov;         // [Nq,M,M]
    double* X;            // per-CTA slots [d,d]
};

template <int FAM>
__global__ void __launch_bounds__(NTH) st_predict_kernel(StPredArgs a) {
    constexpr int n = FamilyDim<FAM>::value;
    extern __shared__ __align__(16) unsigned char smraw[];
    Smem& sm = *reinterpret_cast<Smem*>(smraw);
    const int M = a.M, d = M * n;
    double* X = a.X + (size_t)blockIdx.x * d * d;
    MaternGen<FAM, 1> gen;
    gen.spec = a.spec;
    gen.dt = nullptr;
    for (long long q = blockIdx.x; q < a.Nq; q += gridDim.x) {
        const double xt = a.xs[q];
        long long lo = 0, hi = a.N;  // number of training times strictly below xt = index into the augmented arrays
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (a.x[mid] < xt) lo = mid + 1; else hi = mid;
        }
        const long long ind = lo;
        const double xl = ind == 0 ? -1e10 : a.x[ind - 1];
        const double xr = ind == a.N ? 1e10 : a.x[ind];
        double P1[n * n], W[n * n], Tm[n * n];
        cond_stats(gen, xt - xl, xr - xt, P1, W, Tm);
        const double* Cl = ind == 0 ? a.pk : a.sP + (size_t)(ind - 1) * d * d;
        const double* Cr = ind == a.N ? a.pk : a.sP + (size_t)ind * d * d;
        const double* G = ind == 0 ? nullptr : a.gain + (size_t)(ind - 1) * d * d;
        const double* ml = ind == 0 ? nullptr : a.sm + (size_t)(ind - 1) * d;
        const double* mr = ind == a.N ? nullptr : a.sm + (size_t)ind * d;
        // X = G C_+  (C_+ symmetric: NT form)
        const int tm = (d + 63) / 64;
        for (int t = 0; t < tm * tm; ++t) {
            const int i0 = (t / tm) * 64, j0 = (t % tm) * 64;
            double acc[4][4];
            if (G) tile_nt<64, 64, false>(G, d, d, Cr, d, d, d, i0, j0, nullptr, acc, sm);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int row = i0 + trow<64>(r), col = j0 + tcol<64>(c);
                    if (row < d && col < d) X[(size_t)row * d + col] = G ? acc[r][c] : 0.0;
                }
        }
        __syncthreads();
        __threadfence_block();
        for (int e = threadIdx.x; e < M * M; e += NTH) {
            const int i = e / M, j = e % M;
            double v = (i == j) ? Tm[0] : 0.0;
#pragma unroll
            for (int p = 0; p < n; ++p)
#pragma unroll
                for (int b = 0; b < n; ++b) {
                    const size_t rij = (size_t)(i * n + p) * d + j * n + b, rji = (size_t)(j * n + p) * d + i * n + b;
                    v = fma(P1[p] * P1[b], Cl[rij], v);
                    v = fma(W[p] * W[b], Cr[rij], v);
                    v = fma(P1[p] * W[b], ldg(X + rij) + ldg(X + rji), v);
                }
            a.fcov[(size_t)q * M * M + e] = v;
        }
        for (int i = threadIdx.x; i < M; i += NTH) {
            double v = 0.0;
#pragma unroll
            for (int p = 0; p < n; ++p) v = fma(P1[p], ml ? ml[i * n + p] : 0.0, fma(W[p], mr ? mr[i * n + p] : 0.0, v));
            a.fmean[(size_t)q * M + i] = v;
        }
        __syncthreads();
    }
}

// ---- Gaussian KL term with full M x M blocks (utils.py:510-531 through basemodels.py:715-721) ------------------
// out_k = log N(y_k | m_k, R_k) - 0.5 tr(R_k^-1 V_k), both through chol(R_k).  One CTA per time step.
// T = [R ; V ; I ; (y - m)^T]:  U = V L^-T, LiT = L^-T  =>  tr(R^-1 V) = sum_ij LiT[i][j] U[i][j].
__global__ void __launch_bounds__(NTH) st_gell_kernel(long long N, int n, const double* y, const double* m, const double* V,
                                                      const double* R, const uint8_t* mask, double* out, double* Tall) {
    extern __shared__ __align__(16) unsigned char smraw[];
    Smem& sm = *reinterpret_cast<Smem*>(smraw);
    __shared__ double red[NTH];
    const int np = pad32(n), rows = 3 * np + NB;
    double* T = Tall + (size_t)blockIdx.x * rows * np;
    for (long long k = blockIdx.x; k < N; k += gridDim.x) {
        for (int e = threadIdx.x; e < rows * np; e += NTH) {
            const int r = e / np, c = e % np;
            double v = 0.0;
            // mask rules of utils.py:522-527: masked entries independent, noise variance 1/(2 pi), covariance 1e-20, residual 0
            const uint8_t* mk = mask ? mask + (size_t)k * n : nullptr;
            if (r < np) {
                if (r < n && c < n) {
                    v = R[(size_t)k * n * n + (size_t)r * n + c];
                    if (mk && (mk[r] || mk[c])) v = (r == c) ? 0.15915494309189535 : 0.0;
                } else if (r == c) v = 1.0;
            } else if (r < 2 * np) {
                const int rr = r - np;
                if (rr < n && c < n) {
                    v = V[(size_t)k * n * n + (size_t)rr * n + c];
                    if (mk && (mk[rr] || mk[c])) v = (rr == c) ? 1e-20 : 0.0;
                }
            } else if (r < 3 * np) {
                v = (r - 2 * np == c) ? 1.0 : 0.0;
            } else if (r == 3 * np && c < n) {
                v = (mk && mk[c]) ? 0.0 : y[(size_t)k * n + c] - m[(size_t)k * n + c];
            }
            T[e] = v;
        }
        __syncthreads();
        __threadfence_block();
        chol_stack_cta(T, np, rows, sm, 2 * np / NB);
        const double* U = T + (size_t)np * np;
        const double* LiT = T + (size_t)2 * np * np;
        const double* z = T + (size_t)3 * np * np;
        double s = 0.0;
        for (int e = threadIdx.x; e < n * np; e += NTH) s = fma(ldg(LiT + e), ldg(U + e), s);
        double q = 0.0;
        for (int c = threadIdx.x; c < n; c += NTH) {
            const double zc = ldg(z + c);
            q += zc * zc + 2.0 * log(ldg(T + (size_t)c * np + c));
        }
        red[threadIdx.x] = -0.5 * (q + s);
        __syncthreads();
        for (int off = NTH / 2; off > 0; off >>= 1) {
            if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[k] = red[0] - 0.5 * n * 1.8378770664093453;
        __syncthreads();
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
template <class Kern>
static cudaError_t allow_smem(Kern k) {
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

// The persistent kernels synchronise their CTAs through an L2-atomic grid barrier, which is only safe when every CTA
// of the grid is resident at once: they are launched cooperatively, so a grid that does not fit (another context
// holding SMs, a smaller part) fails at launch with cudaErrorCooperativeLaunchTooLarge instead of hanging.
template <class Kern, class Args>
static cudaError_t launch_persistent(Kern k, int grid, const Args& a, cudaStream_t s) {
    cudaError_t e = allow_smem(k);
    if (e != cudaSuccess) return e;
    void* params[] = {(void*)&a};
    return cudaLaunchCooperativeKernel((const void*)k, dim3((unsigned)grid), dim3(NTH), params, kSmemBytes, s);
}

static cudaError_t zero_prof(cudaStream_t s) {
    void* p = nullptr;
    cudaError_t e = cudaGetSymbolAddress(&p, g_prof);
    return e != cudaSuccess ? e : cudaMemsetAsync(p, 0, sizeof(long long) * 16, s);
}

static int sm_count();
static int sm_count_cached() {
    static int v = 0;
    if (v == 0) v = sm_count();
    return v;
}
static int sm_count() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 1;
}

struct Carver {
    char* p;
    size_t left;
    bool ok = true;
    template <class X>
    X* take(size_t count) {
        size_t bytes = (count * sizeof(X) + 255) / 256 * 256;
        if (bytes > left) { ok = false; return nullptr; }
        X* r = (X*)p;
        p += bytes;
        left -= bytes;
        return r;
    }
};

static size_t filter_ws(int M, int n) {
    size_t d = (size_t)M * n, Mp = pad32(M), dp = pad32((int)d);
    return ((Mp + dp + NB) * Mp + (Mp + NB) * Mp + (2 * Mp + dp + 2 * NB) * NB + 4 * kLpub + d * d + d + 64) * sizeof(double) + 12 * 256;
}
static long long smoother_chunk(long long N, int M, int n) {
    // time steps per launch pair when the caller does not keep the gains: ~2 GB of gain scratch
    const size_t d = (size_t)M * n;
    long long c = (long long)((2ull << 30) / (d * d * sizeof(double)));
    if (c < 2 * (long long)sm_count_cached()) c = 2 * (long long)sm_count_cached();
    return c < N ? c : N;
}
static size_t smoother_ws(int M, int n, long long N, bool own_gains) {
    const size_t d = (size_t)M * n, dp = pad32((int)d);
    const size_t slots = (size_t)(sm_count_cached() < N ? sm_count_cached() : (N > 0 ? N : 1));
    size_t doubles = slots * 3 * dp * dp + 3 * d * d + 2 * d + 64;
    if (own_gains) doubles += (size_t)smoother_chunk(N, M, n) * d * d;
    return doubles * sizeof(double) + 16 * 256;
}
static int batch_grid(long long N) {
    long long g = 2LL * sm_count();
    return (int)(N < g ? (N > 0 ? N : 1) : g);
}
static size_t inverse_ws(long long N, int n) {
    size_t np = pad32(n);
    return (size_t)batch_grid(N) * (2 * np + NB) * np * sizeof(double) + 256;
}
static size_t gell_ws(long long N, int n) {
    size_t np = pad32(n);
    return (size_t)batch_grid(N) * (3 * np + NB) * np * sizeof(double) + 256;
}

}  // namespace st
}  // namespace bn

using namespace bn;
using namespace bn::st;

#define ST_DISPATCH_FAMILY(fam, CALL)                  \
    switch (fam) {                                     \
        case BN_MATERN12: { CALL(BN_MATERN12); break; } \
        case BN_MATERN32: { CALL(BN_MATERN32); break; } \
        case BN_MATERN52: { CALL(BN_MATERN52); break; } \
        default: break;                                \
    }

static int st_check_spec(const bn_kernel_spec* k, int M, int* n_out) {
    BN_REQUIRE(k != nullptr, "temporal kernel spec is null");
    BN_REQUIRE(k->n_components == 1, "the temporal kernel of a spatio-temporal prior has one component");
    BN_REQUIRE(k->family == BN_MATERN12 || k->family == BN_MATERN32 || k->family == BN_MATERN52,
               "temporal family %d not available on the dense path (Matern-1/2, -3/2, -5/2)", k->family);
    BN_REQUIRE(M >= 1 && M <= 4096, "M = %d spatial points out of range", M);
    *n_out = family_dim(k->family);
    return 0;
}

static size_t mf_filter_ws(int M);
extern "C" size_t bn_st_workspace_bytes(const bn_kernel_spec* temporal, int M, int64_t N, int Ns) {
    int n = 0;
    if (st_check_spec(temporal, M, &n) != 0) return 0;
    size_t a = filter_ws(M, n), b = smoother_ws(M, n, N, true), c = inverse_ws(N, M), e = gell_ws(N, M);
    (void)Ns;
    size_t m = a > b ? a : b;
    if (mf_filter_ws(M) > m) m = mf_filter_ws(M);
    if (c > m) m = c;
    if (e > m) m = e;
    return m;
}

extern "C" int bn_st_kalman_filter(const bn_kernel_spec* temporal, int M, int64_t N, const double* dt, const double* y,
                                   const double* noise_cov, const uint8_t* mask, int return_predict, double* ell,
                                   double* means, double* covs, void* workspace, size_t workspace_bytes, void* stream) {
    int n = 0;
    if (int rc = st_check_spec(temporal, M, &n)) return rc;
    BN_REQUIRE(N >= 0, "N must be non-negative");
    if (N == 0) return 0;
    BN_REQUIRE(dt && y && noise_cov && means && covs, "null array");
    BN_REQUIRE(workspace && workspace_bytes >= filter_ws(M, n), "workspace too small: %zu bytes needed", filter_ws(M, n));
    const size_t d = (size_t)M * n, Mp = pad32(M), dp = pad32((int)d);
    Carver cv{(char*)workspace, workspace_bytes};
    FilterArgs a;
    a.spec = *temporal; a.M = M; a.N = N; a.dt = dt; a.y = y; a.R = noise_cov; a.mask = mask; a.return_predict = return_predict;
    a.ell = ell; a.means = means; a.covs = covs;
    a.ctr = cv.take<unsigned long long>(32);
    a.ellacc = cv.take<double>(32);
    a.T = cv.take<double>((Mp + dp + NB) * Mp);
    a.T2 = cv.take<double>((Mp + NB) * Mp);
    a.Pacc = cv.take<double>(((Mp + dp + NB) / NB + (Mp + NB) / NB) * NB * NB);
    a.Lpub = cv.take<double>(4 * kLpub);
    a.Pcur = cv.take<double>(d * d);
    a.mcur = cv.take<double>(d);
    BN_REQUIRE(cv.ok, "workspace carve failed");
    cudaStream_t s = (cudaStream_t)stream;
    BN_CUDA(cudaMemsetAsync(a.ctr, 0, 256, s));
    BN_CUDA(zero_prof(s));
    const int grid = sm_count();
    BN_REQUIRE(grid >= 2, "the dense path needs at least 2 SMs");
#define CALL(F) BN_LAUNCH("st_filter", s, BN_CUDA(launch_persistent(st_filter_kernel<F>, grid, a, s)))
    ST_DISPATCH_FAMILY(temporal->family, CALL)
#undef CALL
    BN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int bn_st_rts_smoother(const bn_kernel_spec* temporal, int M, int64_t N, const double* dt,
                                  const double* filter_mean, const double* filter_cov, int return_full, double* means,
                                  double* covs, double* gains, void* workspace, size_t workspace_bytes, void* stream) {
    int n = 0;
    if (int rc = st_check_spec(temporal, M, &n)) return rc;
    BN_REQUIRE(N >= 0, "N must be non-negative");
    if (N == 0) return 0;
    BN_REQUIRE(dt && filter_mean && filter_cov && means && covs, "null array");
    const size_t need = smoother_ws(M, n, N, gains == nullptr);
    BN_REQUIRE(workspace && workspace_bytes >= need, "workspace too small: %zu bytes needed", need);
    const size_t d = (size_t)M * n, dp = pad32((int)d);
    const long long chunk = gains ? (long long)N : smoother_chunk(N, M, n);
    const int grid = sm_count_cached();
    const int ggrid = (int)((long long)grid < N ? (long long)grid : N);
    Carver cv{(char*)workspace, workspace_bytes};
    GainArgs g;
    g.spec = *temporal; g.M = M; g.dt = dt; g.fP = filter_cov;
    SmootherArgs a;
    a.spec = *temporal; a.M = M; a.N = N; a.dt = dt; a.fm = filter_mean; a.fP = filter_cov; a.return_full = return_full;
    a.means = means; a.covs = covs;
    a.ctr = cv.take<unsigned long long>(32);
    g.T = cv.take<double>((size_t)ggrid * 3 * dp * dp);
    a.sP = cv.take<double>(d * d);
    a.smv = cv.take<double>(d);
    a.Dm = cv.take<double>(d * d);
    a.v = cv.take<double>(d);
    a.Z = cv.take<double>(d * d);
    double* gbuf = gains ? gains : cv.take<double>((size_t)chunk * d * d);
    BN_REQUIRE(cv.ok, "workspace carve failed");
    cudaStream_t s = (cudaStream_t)stream;
    BN_CUDA(zero_prof(s));
    for (long long k1 = N; k1 > 0; k1 -= chunk) {
        const long long k0 = k1 - chunk > 0 ? k1 - chunk : 0;
        g.k0 = k0; g.k1 = k1; g.G = gains ? gains + (size_t)k0 * d * d : gbuf;
        a.k0 = k0; a.k1 = k1; a.G = g.G; a.first = (k1 == N);
        BN_CUDA(cudaMemsetAsync(a.ctr, 0, 256, s));
#define CALL(F) BN_CUDA(allow_smem(st_gain_kernel<F>)); BN_LAUNCH("st_gain", s, st_gain_kernel<F><<<ggrid, NTH, kSmemBytes, s>>>(g)); \
                BN_LAUNCH("st_smoother", s, BN_CUDA(launch_persistent(st_smoother_kernel<F>, grid, a, s)))
        ST_DISPATCH_FAMILY(temporal->family, CALL)
#undef CALL
        BN_CUDA(cudaGetLastError());
    }
    return 0;
}

static size_t mf_filter_ws(int M) {
    size_t Mp = pad32(M);
    return ((2 * Mp + NB) * Mp + (Mp + NB) * Mp + (3 * Mp + 2 * NB) * NB + 4 * kLpub + 64) * sizeof(double) + 12 * 256;
}

extern "C" int bn_st_kalman_filter_meanfield(const bn_kernel_spec* temporal, int M, int64_t N, const double* dt,
                                             const double* y, const double* noise_cov, const uint8_t* mask, double* ell,
                                             double* means, double* covs, void* workspace, size_t workspace_bytes,
                                             void* stream) {
    int n = 0;
    if (int rc = st_check_spec(temporal, M, &n)) return rc;
    BN_REQUIRE(N >= 0, "N must be non-negative");
    if (N == 0) return 0;
    BN_REQUIRE(dt && y && noise_cov && means && covs, "null array");
    BN_REQUIRE(workspace && workspace_bytes >= mf_filter_ws(M), "workspace too small: %zu bytes needed", mf_filter_ws(M));
    const size_t Mp = pad32(M);
    Carver cv{(char*)workspace, workspace_bytes};
    MfFilterArgs a;
    a.spec = *temporal; a.M = M; a.N = N; a.dt = dt; a.y = y; a.R = noise_cov; a.mask = mask; a.ell = ell;
    a.means = means; a.covs = covs;
    a.ctr = cv.take<unsigned long long>(32);
    a.T = cv.take<double>((2 * Mp + NB) * Mp);
    a.T2 = cv.take<double>((Mp + NB) * Mp);
    a.Pacc = cv.take<double>(((2 * Mp + NB) / NB + (Mp + NB) / NB) * NB * NB);
    a.Lpub = cv.take<double>(4 * kLpub);
    BN_REQUIRE(cv.ok, "workspace carve failed");
    cudaStream_t s = (cudaStream_t)stream;
    BN_CUDA(cudaMemsetAsync(a.ctr, 0, 256, s));
    BN_CUDA(zero_prof(s));
    const int grid = sm_count();
    BN_REQUIRE(grid >= 2, "the dense path needs at least 2 SMs");
#define CALL(F) BN_LAUNCH("st_mf_filter", s, BN_CUDA(launch_persistent(st_mf_filter_kernel<F>, grid, a, s)))
    ST_DISPATCH_FAMILY(temporal->family, CALL)
#undef CALL
    BN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int bn_st_rts_smoother_meanfield(const bn_kernel_spec* temporal, int M, int64_t N, const double* dt,
                                            const double* filter_mean, const double* filter_cov, int return_full,
                                            double* means, double* covs, double* gains, void* stream) {
    int n = 0;
    if (int rc = st_check_spec(temporal, M, &n)) return rc;
    BN_REQUIRE(N >= 0, "N must be non-negative");
    if (N == 0) return 0;
    BN_REQUIRE(dt && filter_mean && filter_cov && means && covs, "null array");
    cudaStream_t s = (cudaStream_t)stream;
    if (!return_full) BN_CUDA(cudaMemsetAsync(covs, 0, (size_t)N * M * M * sizeof(double), s));
    const unsigned grid = (unsigned)((M + 63) / 64);
#define CALL(F) BN_LAUNCH("st_mf_smoother", s, st_mf_smoother_kernel<F><<<grid, 64, 0, s>>>(*temporal, M, N, dt, filter_mean, filter_cov, return_full, means, covs, gains))
    ST_DISPATCH_FAMILY(temporal->family, CALL)
#undef CALL
    BN_CUDA(cudaGetLastError());
    return 0;
}


static __global__ void st_pinf_kron_kernel(bn_kernel_spec spec, int M, int n, double* pk) {
    // I (x) Pinf_t, d x d
    const int d = M * n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)d * d; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / d), c = (int)(e % d);
        double v = 0.0;
        if (r / n == c / n) {
            double Pp[6];
            const int a = r % n, b = c % n;
            if (spec.family == BN_MATERN12) MaternBlock<BN_MATERN12, double>::pinf(spec.variance[0], spec.lengthscale[0], Pp);
            else if (spec.family == BN_MATERN32) MaternBlock<BN_MATERN32, double>::pinf(spec.variance[0], spec.lengthscale[0], Pp);
            else MaternBlock<BN_MATERN52, double>::pinf(spec.variance[0], spec.lengthscale[0], Pp);
            v = Pp[sidx(a, b)];
        }
        pk[e] = v;
    }
}

extern "C" size_t bn_st_predict_workspace_bytes(const bn_kernel_spec* temporal, int M, int64_t Nq) {
    int n = 0;
    if (st_check_spec(temporal, M, &n) != 0) return 0;
    const size_t d = (size_t)M * n;
    return ((size_t)batch_grid(Nq) + 1) * d * d * sizeof(double) + 512;
}

extern "C" int bn_st_predict_state(const bn_kernel_spec* temporal, int M, int64_t N, const double* x, int64_t Nq,
                                   const double* x_test, const double* mean, const double* cov, const double* gain,
                                   double* f_mean, double* f_cov, void* workspace, size_t workspace_bytes, void* stream) {
    int n = 0;
    if (int rc = st_check_spec(temporal, M, &n)) return rc;
    BN_REQUIRE(N >= 1 && Nq >= 0, "bad sizes N = %lld, N_test = %lld", (long long)N, (long long)Nq);
    if (Nq == 0) return 0;
    BN_REQUIRE(x && x_test && mean && cov && gain && f_mean && f_cov, "null array");
    BN_REQUIRE(workspace && workspace_bytes >= bn_st_predict_workspace_bytes(temporal, M, Nq), "workspace too small");
    const size_t d = (size_t)M * n;
    StPredArgs a;
    a.spec = *temporal; a.M = M; a.N = N; a.Nq = Nq; a.x = x; a.xs = x_test; a.sm = mean; a.sP = cov; a.gain = gain;
    a.pk = (double*)workspace; a.X = (double*)workspace + d * d; a.fmean = f_mean; a.fcov = f_cov;
    cudaStream_t s = (cudaStream_t)stream;
    st_pinf_kron_kernel<<<64, 256, 0, s>>>(*temporal, M, n, (double*)workspace);
    BN_CUDA(cudaGetLastError());
#define CALL(F) BN_CUDA(allow_smem(st_predict_kernel<F>)); BN_LAUNCH("st_predict", s, st_predict_kernel<F><<<batch_grid(Nq), NTH, kSmemBytes, s>>>(a))
    ST_DISPATCH_FAMILY(temporal->family, CALL)
#undef CALL
    BN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int bn_spd_inverse_batched(int64_t N, int n, const double* A, const double* rhs, double jitter, double* inv,
                                      double* sol, double* logdet, void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(N >= 0 && n >= 1 && n <= 4096, "bad sizes N = %lld, n = %d", (long long)N, n);
    if (N == 0) return 0;
    BN_REQUIRE(A && inv, "null array");
    BN_REQUIRE(sol == nullptr || rhs != nullptr, "sol needs rhs");
    BN_REQUIRE(workspace && workspace_bytes >= inverse_ws(N, n), "workspace too small: %zu bytes needed", inverse_ws(N, n));
    InvArgs a{};
    a.N = N; a.n = n; a.Ns = 0; a.A = A; a.rhs_s = rhs; a.jitter = jitter; a.S = nullptr; a.inv = inv; a.sol = sol;
    a.logdet = logdet; a.T = (double*)workspace;
    cudaStream_t s = (cudaStream_t)stream;
    BN_CUDA(allow_smem(st_inverse_kernel<false>));
    BN_LAUNCH("st_inverse", s, st_inverse_kernel<false><<<batch_grid(N), NTH, kSmemBytes, s>>>(a));
    BN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int bn_st_pseudo_lik(int64_t N, int Ns, int M, const double* Bt, const double* nat1, const double* nat2_diag,
                                double jitter, double* pseudo_y, double* pseudo_var, double* nat2_full, double* logdet,
                                void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(N >= 0 && M >= 1 && M <= 4096 && Ns >= 1, "bad sizes N = %lld, Ns = %d, M = %d", (long long)N, Ns, M);
    if (N == 0) return 0;
    BN_REQUIRE(Bt && nat1 && nat2_diag && pseudo_y && pseudo_var, "null array");
    BN_REQUIRE(workspace && workspace_bytes >= inverse_ws(N, M), "workspace too small: %zu bytes needed", inverse_ws(N, M));
    InvArgs a{};
    a.N = N; a.n = M; a.Ns = Ns; a.Bt = Bt; a.lam = nat2_diag; a.rhs_s = nat1; a.jitter = jitter; a.S = nat2_full;
    a.inv = pseudo_var; a.sol = pseudo_y; a.logdet = logdet; a.T = (double*)workspace;
    cudaStream_t s = (cudaStream_t)stream;
    BN_CUDA(allow_smem(st_inverse_kernel<true>));
    BN_LAUNCH("st_pseudo_lik", s, st_inverse_kernel<true><<<batch_grid(N), NTH, kSmemBytes, s>>>(a));
    BN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int bn_st_posterior_to_data(int64_t N, int Ns, int M, const double* B, const double* cdiag,
                                       const double* post_mean, const double* post_cov, double* mean_f, double* var_f,
                                       void* stream) {
    BN_REQUIRE(N >= 0 && M >= 1 && Ns >= 1, "bad sizes");
    if (N == 0) return 0;
    BN_REQUIRE(B && post_mean && post_cov && mean_f && var_f, "null array");
    cudaStream_t s = (cudaStream_t)stream;
    BN_CUDA(allow_smem(st_to_data_kernel));
    BN_LAUNCH("st_to_data", s, st_to_data_kernel<<<batch_grid(N), NTH, kSmemBytes, s>>>(N, Ns, M, B, cdiag, post_mean, post_cov, mean_f, var_f));
    BN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int bn_st_gaussian_expected_log_lik(int64_t N, int M, const double* pseudo_y, const double* post_mean,
                                               const double* post_cov, const double* pseudo_var, const uint8_t* mask, double* values,
                                               double* sum, void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(N >= 0 && M >= 1 && M <= 4096, "bad sizes");
    if (N == 0) return 0;
    BN_REQUIRE(pseudo_y && post_mean && post_cov && pseudo_var && values, "null array (values[N] is required)");
    BN_REQUIRE(workspace && workspace_bytes >= gell_ws(N, M), "workspace too small: %zu bytes needed", gell_ws(N, M));
    cudaStream_t s = (cudaStream_t)stream;
    BN_CUDA(allow_smem(st_gell_kernel));
    BN_LAUNCH("st_gell", s, st_gell_kernel<<<batch_grid(N), NTH, kSmemBytes, s>>>(N, M, pseudo_y, post_mean, post_cov, pseudo_var, mask, values, (double*)workspace));
    BN_CUDA(cudaGetLastError());
    if (sum) {
        BN_LAUNCH("sum", s, sum_kernel<false><<<1, 1024, 0, s>>>(values, N, sum, 1.0));
        BN_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int bn_st_profile(int64_t* cycles_host, int n) {
    BN_REQUIRE(cycles_host && n >= 1 && n <= 16, "bad arguments");
    long long tmp[16];
    BN_CUDA(cudaMemcpyFromSymbol(tmp, g_prof, sizeof(tmp)));  // synchronises with the default stream's work
    for (int i = 0; i < n; ++i) cycles_host[i] = tmp[i];
    return 0;
}
