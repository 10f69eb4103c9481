// oc_hjb_fused.cuh -- stage-fused RK45 step kernel (temporal blocking across the 6 new RHS evaluations).
//
// One launch performs a whole Dormand-Prince step attempt (scipy rk.py:14-71 rk_step + :146-147 error
// estimate) and the dense-output samples phi(t_e) of the t_eval points that fall inside the step
// (rk.py:178-180,723-737), reading y, f = K[0] and coef once and writing y_new, f_new = K[6] and the
// phi slices once: 40 B/cell/attempt + 8 B/cell/emitted slice instead of the 312 B/cell/attempt of the
// stage-wise formulation.
//
// Scheme ("2.5-D" streaming): a CTA owns BX = 256 consecutive columns and marches down a chunk of rows.
// Thread t owns ONE column; all RK state of that column lives in its registers as short row windows.
// At iteration r the CTA loads row r (stage 1 input), and stage s = 2..7 evaluates the stencil on row
// r-(s-1), i.e. every stage lags one row behind the previous one, so the 3-row stencil windows of all
// six stage inputs are register-resident.  Only the left/right neighbours cross threads: each stage
// input row is published once to a double-buffered shared-memory line (one __syncthreads per row).
// Rows of y, f, coef are staged ahead of use through an 8-deep cp.async ring in shared memory.
// Six stencil applications consume HALO = 6 columns per side (8 allocated for 64-byte aligned rows) and
// 6 rows above/below the chunk, which are recomputed redundantly (about 7 % + 5 % extra work).
//
// Arithmetic: FP64, FMA contraction allowed.  The stage combination is evaluated as
// y + (h a_s1) k_1 + (h a_s2) k_2 + ... with premultiplied coefficients; it differs from scipy's
// (sum_j a_sj k_j) * h by rounding only (parity bar 1e-10, observed ~1e-13).
#pragma once

namespace fused {

#ifndef OC_BX
#define OC_BX 128
#endif
#ifndef OC_PF
#define OC_PF 6
#endif
#ifndef OC_PD
#define OC_PD (OC_PF - 1)
#endif
constexpr int BX = OC_BX;           // threads per CTA = columns per tile incl. halo
constexpr int HX = 8;               // halo columns per side (6 needed)
constexpr int VX = BX - 2 * HX;     // valid output columns per tile (112)
constexpr int HY = 6;               // halo rows per side
constexpr int PF = OC_PF;           // cp.async ring depth (rows): rows r .. r+PD
constexpr int PD = OC_PD;           // prefetch distance (rows in flight)
constexpr int UNROLL = 12;          // rows per unrolled loop body (multiple of PF and of 2)
static_assert(UNROLL % PF == 0 && UNROLL % 2 == 0 && PD < PF, "ring / unroll geometry");
#ifndef OC_CTAS
#define OC_CTAS 2
#endif
#if !defined(OC_BRANCHY) && !defined(OC_BRANCHLESS)
#define OC_BRANCHLESS 1
#endif
#if !defined(OC_VREGCONST) && !defined(OC_LDSCONST) && defined(OC_BRANCHLESS)
#define OC_VREGCONST 3
#endif
// 2 CTAs (8 warps) per SM with 192 registers/thread beat 3 CTAs at 164: the row body is one long straight-line block
// and the extra registers buy instruction-level parallelism across the six stage chains (measured 0.75 vs 0.70 of
// the HBM peak, profiles/r1_fused_v4.md)
constexpr int CTAS_PER_SM = OC_CTAS;
constexpr int NE_MAX = 6;           // dense-output samples per launch

// rows per CTA chunk: balance the tail of the last wave (slots = SMs x resident CTAs) against the 2*HY halo rows
// each chunk recomputes.  `rows` = rows of the band, `copies` = bands launched back to back on this GPU.
inline int plan_chunk_rows(int Nx, int rows, int n_sm, int copies, int must_divide) {
    static const int cand[] = {16, 32, 48, 64, 96, 128, 160, 192, 224, 256, 320, 384, 448, 512, 640, 768, 1024, 2048};
    const long long slots = (long long)n_sm * CTAS_PER_SM;
    const int gx = (Nx + VX - 1) / VX;
    double best = 1e300;
    int rc = 16;
    for (int c : cand) {
        if (must_divide && rows % c) continue;
        const int cc = c < rows ? c : rows;
        const long long blocks = (long long)gx * ((rows + cc - 1) / cc) * copies;
        const double cost = (double)((blocks + slots - 1) / slots) * (cc + 2 * HY);
        if (cost < best) { best = cost; rc = cc; }
    }
    return rc;
}

constexpr int P2P_MAX_RANKS = 8;    // one NVSwitch box
constexpr int P2P_INBOX_STRIDE = 64; // doubles per (parity, rank) inbox slot: <= 62 chunk-row sums + the sequence number

struct Args {
    const double *y, *k1, *coef;
    double *ynew, *k7, *partial;
    double *phi[NE_MAX];
    double ha21;
    double ha3[2], ha4[3], ha5[4], ha6[5];
    double hb[6];            // h*B (B[1] = 0 unused)
    double he[7];            // h*E (E[1] = 0 unused)
    double w[NE_MAX][7];     // h * sum_q P[j][q] x_e^(q+1)
    double A;                // -0.5 sigma^2 / (dx dy)
    double rtol, atol;
    int Ny, Nx, RC;          // global grid, RC = rows per chunk
    // row band (multi-GPU / virtual ranks): rows [own0, own1) are computed; y, k1, coef, ynew, k7 point to storage
    // whose first row is global row `row_base` (halo rows from the neighbouring bands included); phi slices
    // start at global row `phi_row_base`.  Single GPU: row_base = phi_row_base = 0, own = [0, Ny).
    int row_base, own0, own1, phi_row_base;
    // Final reduction inside the launch (single-GPU and batched solves): the CTA that finishes last adds the per-CTA
    // partial sums in the library's fixed order (per chunk row: 256 strided accumulators, shuffle tree, 8 warp sums left
    // to right; then the chunk rows top to bottom -- exactly rowgroup_sum_kernel + the host loop) and publishes the
    // scalar in page-locked host memory, where the host polls for it: no reduction launch, no copy, no stream
    // synchronisation per step attempt.  ticket == NULL: partial sums only (row-band solver: NCCL all-gather).
    unsigned *ticket;
    double *result;                 // mapped host memory: result[0] = sum
    unsigned long long *result_seq; // mapped host memory: set to `seq` after result[0] is visible
    unsigned long long seq;
    // Row-band solve over NVLink peer memory (template parameter P2P; oc_hjb_dist.cu): the halo exchange and the
    // error-norm all-gather are part of this launch.  (a) A thread that stores a row of y_new / f_new lying within
    // HY rows of a band edge stores it a second time, straight into the neighbouring GPU's halo rows.  (b) The last
    // CTA writes this band's chunk-row sums into the inbox of every rank (peer stores, then a system fence, then the
    // sequence number), waits for the inboxes of all ranks, adds the sums in global chunk order and hands the total to
    // the host.  A rank that has seen every rank's sequence number also knows that every rank's kernel -- and with it
    // every halo store aimed at this rank -- has completed.
    double *peer_ynew[2], *peer_k7[2];  // [0]: rank-1 (rows above), [1]: rank+1; NULL at the ends of the grid
    int peer_row_base[2];               // global row of storage row 0 of the neighbour's arrays
    int nranks, my_rank;
    double *peer_inbox[P2P_MAX_RANKS];  // inbox of rank q as seen from this GPU; [my_rank] is the local one
};
constexpr int MAX_FINAL_ROWS = 2 * 6 * (BX + 2) - 8;  // chunk-row sums are staged in the (then idle) exchange buffer

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
// predicated forms: the unrolled row body has no divergent region (a branch around the loads / stores makes ptxas save
// and restore the uniform registers that hold the RK coefficients on both sides of it)
__device__ __forceinline__ void cp_async8_if(void *smem, const void *gmem, bool ok) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p cp.async.ca.shared.global [%0], [%1], 8; }"
                 ::"r"(s), "l"(gmem), "r"((int)ok) : "memory");
}
__device__ __forceinline__ void st_if(double *p, double v, bool ok) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p st.global.f64 [%0], %1; }" ::"l"(p), "d"(v), "r"((int)ok) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// optimals.py:154-162.  Mirror ghosts (:149-152) need no special case here: rows/columns outside the grid
// are LOADED from their mirror image (row -j -> row j, row Ny-1+j -> Ny-1-j), the stencil commutes with
// that reflection, so every stage input is automatically even about the boundary and the value the
// reference reads from its ghost cell (ghost(-1) = value(1)) is exactly what sits in the halo.
// The newest row `dn` of the stage input is produced in the same iteration, so the evaluation is split
// into a part that does not depend on it (`pre`) and one FMA that does: the serial chain through the six
// stages of one row iteration is 2 FMAs per stage.
__device__ __forceinline__ double rhs_pre(double up, double lf, double rt, double C, double cf, double A) {
    double S = (up + lf) + fma(-4.0, C, rt);
    return fma(A, S, -(cf * C));
}
__device__ __forceinline__ double rhs_fin(double pre, double dn, double cf, double A) {
    double r = fma(A, dn, pre);
#ifdef OC_WALLINT
    return (__double2hiint(cf) == 0x7ff80000) ? 0.0 : r;  // wall marker written by prep_kernel (integer pipe)
#else
    return (cf != cf) ? 0.0 : r;  // wall: optimals.py:162
#endif
}

// 1/x for finite x > 0 (the error scale is >= atol): hardware seed + two Newton steps (~1e-16), no slow path
__device__ __forceinline__ double rcp_pos(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

struct Smem {
    double ex[2][6][BX + 2];   // published stage-input rows (u2..u6, y_new), double buffered
    double pf[PF][3][BX];      // cp.async ring: y, k1, coef
    double red[BX / 32];
#if defined(OC_LDSCONST) || defined(OC_VREGCONST)
    double cst[(1 + NE_MAX) * 6];  // h*E and the dense-output weights, read as broadcast LDS.128 (frees uniform registers)
#endif
};

__device__ __forceinline__ int mirror(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

template <int NE, bool P2P = false>
__global__ void __launch_bounds__(BX, CTAS_PER_SM) hjb_fused_kernel(const Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x;
    const int gx = blockIdx.x * VX - HX + tid;
    const int gxm = min(max(mirror(gx, a.Nx), 0), a.Nx - 1);  // column this thread loads (clamped far outside)
    const int y0 = a.own0 + blockIdx.y * a.RC;
    const int y1 = min(y0 + a.RC, a.own1);
    const bool col_out = tid >= HX && tid < BX - HX && gx < a.Nx;  // columns this thread stores
    const int phi_shift = (a.row_base - a.phi_row_base) * a.Nx;    // phi slices may start at another global row

    for (int i = tid; i < 2 * 6 * (BX + 2); i += BX) (&sm.ex[0][0][0])[i] = 0.0;
#if defined(OC_LDSCONST) || defined(OC_VREGCONST)
    if (tid < 6) {
        const int j = tid == 0 ? 0 : tid + 1;  // stages 1,3,4,5,6,7 (B[1] = E[1] = P[1] = 0)
        sm.cst[tid] = a.he[j];
#pragma unroll
        for (int ee = 0; ee < NE; ee++) sm.cst[6 * (ee + 1) + tid] = a.w[ee][j];
    }
#endif

    // windows, indexed by lag (row r - lag)
    // all windows are registers.  Shared memory is the bottleneck resource of this kernel (every LDS.64 of a
    // warp costs 2 cycles of the SM's 128 B/cycle), so each loaded value is read from the ring exactly once.
    double yw[7], k1w[7], cw[7], k2w[5], k3w[7], k4w[7], k5w[7], k6w[7];
    double u2[3], u3[4], u4[5], u5[6], u6[7], un[8];
#pragma unroll
    for (int i = 0; i < 7; i++) { yw[i] = 0; k1w[i] = 0; cw[i] = 0; k3w[i] = 0; k4w[i] = 0; k5w[i] = 0; k6w[i] = 0; }
#pragma unroll
    for (int i = 0; i < 5; i++) k2w[i] = 0;
    u2[0] = u2[1] = u2[2] = 0; u3[1] = u3[2] = u3[3] = 0; u4[2] = u4[3] = u4[4] = 0;
    u5[3] = u5[4] = u5[5] = 0; u6[4] = u6[5] = u6[6] = 0; un[5] = un[6] = un[7] = 0;

    const int r_begin = y0 - HY, r_end = y1 + HY;  // rows loaded: [r_begin, r_end)
    // ring slot of row `row` = (row - r_begin) mod PF, tracked incrementally (PF is not a power of two)
    auto issue = [&](int row, int slot) {
#ifdef OC_BRANCHLESS
        {
            const int rm = min(max(mirror(row, a.Ny), 0), a.Ny - 1);
            const int g = (rm - a.row_base) * a.Nx + gxm;
            const bool ok = row < r_end;
            cp_async8_if(&sm.pf[slot][0][tid], a.y + g, ok);
            cp_async8_if(&sm.pf[slot][1][tid], a.k1 + g, ok);
            cp_async8_if(&sm.pf[slot][2][tid], a.coef + g, ok);
        }
        if (false) {
#else
        if (row < r_end) {
#endif
            // element offsets fit 32 bits (one field < 2^31 elements); one IMAD instead of 64-bit multiplies
            const int rm = min(max(mirror(row, a.Ny), 0), a.Ny - 1);
            const int g = (rm - a.row_base) * a.Nx + gxm;
            cp_async8(&sm.pf[slot][0][tid], a.y + g);
            cp_async8(&sm.pf[slot][1][tid], a.k1 + g);
            cp_async8(&sm.pf[slot][2][tid], a.coef + g);
        }
        cp_async_commit();  // one group per iteration, possibly empty
    };

#pragma unroll 1
    for (int q = 0; q < PD; q++) issue(r_begin + q, q);
    __syncthreads();

#ifdef OC_VREGCONST
    // The ~47 FP64 constants of the body do not fit the uniform register file (ptxas then spills uniform registers
    // through MOV.SPILL / R2UR.FILL every row).  With 2 CTAs/SM there are spare vector registers: the dense-output
    // weights of the first OC_VREGCONST samples (and h*E) are pinned there by making them opaque to the compiler.
    constexpr int NV = NE < OC_VREGCONST ? NE : OC_VREGCONST;
    double wv[NV > 0 ? NV : 1][6], hev[6];
    // (read back from shared memory: a value ptxas could re-derive from the constant bank would not stay in a register)
    __syncthreads();
#pragma unroll
    for (int ee = 0; ee < NV; ee++)
#pragma unroll
        for (int j = 0; j < 6; j++) wv[ee][j] = sm.cst[6 * (ee + 1) + j];
#ifdef OC_VREGCONST_E
#pragma unroll
    for (int j = 0; j < 6; j++) hev[j] = sm.cst[j];
#else
#pragma unroll
    for (int j = 0; j < 6; j++) hev[j] = a.he[j == 0 ? 0 : j + 1];
#endif
#endif
    double acc = 0.0;
    // The row loop is unrolled by the ring depth: inside the unrolled body the ring slot of every window row, the
    // exchange buffer and the register holding each window entry are compile-time constants -- no address
    // arithmetic and no register moves for the shifting windows.
#pragma unroll 1
    for (int rb = r_begin; rb < r_end; rb += UNROLL) {
#pragma unroll
      for (int uu = 0; uu < UNROLL; uu++) {
        const int r = rb + uu;
        // rows past r_end are harmless (no loads are issued, every store is guarded): no early exit, so the
        // unrolled body is one straight-line block and the window shifts below are pure register renaming
        const int s0 = uu % PF;      // ring slot of row r: (r - r_begin) mod PF
        const int buf = uu & 1;
        issue(r + PD, (s0 + PD) % PF);
        cp_async_wait<PD>();
        yw[0] = sm.pf[s0][0][tid];
        k1w[0] = sm.pf[s0][1][tid];
        cw[0] = sm.pf[s0][2][tid];
        const double *exp_ = &sm.ex[buf ^ 1][0][tid];  // previous iteration's rows: [s*(BX+2)] left, [+2] right
        double *exc = &sm.ex[buf][0][tid + 1];
        // parts of the six stencils that do not depend on this iteration's new rows
        const double p2 = rhs_pre(u2[2], exp_[0 * (BX + 2)], exp_[0 * (BX + 2) + 2], u2[1], cw[1], a.A);
        const double p3 = rhs_pre(u3[3], exp_[1 * (BX + 2)], exp_[1 * (BX + 2) + 2], u3[2], cw[2], a.A);
        const double p4 = rhs_pre(u4[4], exp_[2 * (BX + 2)], exp_[2 * (BX + 2) + 2], u4[3], cw[3], a.A);
        const double p5 = rhs_pre(u5[5], exp_[3 * (BX + 2)], exp_[3 * (BX + 2) + 2], u5[4], cw[4], a.A);
        const double p6 = rhs_pre(u6[6], exp_[4 * (BX + 2)], exp_[4 * (BX + 2) + 2], u6[5], cw[5], a.A);
        const double p7 = rhs_pre(un[7], exp_[5 * (BX + 2)], exp_[5 * (BX + 2) + 2], un[6], cw[6], a.A);
        // stage-input partial sums that do not depend on this iteration's new k's (rk.py:63-64, premultiplied by h)
        const double q3 = fma(a.ha3[0], k1w[1], yw[1]);
        const double q4 = fma(a.ha4[1], k2w[2], fma(a.ha4[0], k1w[2], yw[2]));
        const double q5 = fma(a.ha5[2], k3w[3], fma(a.ha5[1], k2w[3], fma(a.ha5[0], k1w[3], yw[3])));
        const double q6 = fma(a.ha6[3], k4w[4], fma(a.ha6[2], k3w[4], fma(a.ha6[1], k2w[4], fma(a.ha6[0], k1w[4], yw[4]))));
        const double qn = fma(a.hb[4], k5w[5], fma(a.hb[3], k4w[5], fma(a.hb[2], k3w[5], fma(a.hb[0], k1w[5], yw[5]))));
        // ---- the serial chain of this row iteration
        u2[0] = fma(a.ha21, k1w[0], yw[0]);            // stage 2 input on row r
        k2w[1] = rhs_fin(p2, u2[0], cw[1], a.A);       // k2 on row r-1
        u3[1] = fma(a.ha3[1], k2w[1], q3);
        k3w[2] = rhs_fin(p3, u3[1], cw[2], a.A);       // k3 on row r-2
        u4[2] = fma(a.ha4[2], k3w[2], q4);
        k4w[3] = rhs_fin(p4, u4[2], cw[3], a.A);       // k4 on row r-3
        u5[3] = fma(a.ha5[3], k4w[3], q5);
        k5w[4] = rhs_fin(p5, u5[3], cw[4], a.A);       // k5 on row r-4
        u6[4] = fma(a.ha6[4], k5w[4], q6);
        k6w[5] = rhs_fin(p6, u6[4], cw[5], a.A);       // k6 on row r-5
        un[5] = fma(a.hb[5], k6w[5], qn);              // y_new on row r-5 (rk.py:66; B[1] = 0)
        const double k7 = rhs_fin(p7, un[5], cw[6], a.A);  // k7 = f(y_new) on row r-6
        exc[0 * (BX + 2)] = u2[0];
        exc[1 * (BX + 2)] = u3[1];
        exc[2 * (BX + 2)] = u4[2];
        exc[3 * (BX + 2)] = u5[3];
        exc[4 * (BX + 2)] = u6[4];
        exc[5 * (BX + 2)] = un[5];
        // ---- outputs on row r-6
#if defined(OC_LDSCONST)
        {
            const int row = r - 6;
            const bool ok = col_out && row >= y0 && row < y1;
            const int g = (row - a.row_base) * a.Nx + gx;
            const double2 *cs = reinterpret_cast<const double2 *>(sm.cst);
            const double2 e01 = cs[0], e23 = cs[1], e45 = cs[2];
            double e = fma(e45.y, k7, fma(e45.x, k6w[6], fma(e23.y, k5w[6], fma(e23.x, k4w[6], fma(e01.y, k3w[6], e01.x * k1w[6])))));
            double sc = fma(fmax(fabs(yw[6]), fabs(un[6])), a.rtol, a.atol);
            double qq = e * rcp_pos(sc);
            acc = ok ? fma(qq, qq, acc) : acc;
            st_if(a.ynew + g, un[6], ok);
            st_if(a.k7 + g, k7, ok);
#pragma unroll
            for (int ee = 0; ee < NE; ee++) {
                const double2 w01 = cs[3 * (ee + 1)], w23 = cs[3 * (ee + 1) + 1], w45 = cs[3 * (ee + 1) + 2];
                double ph = fma(w45.y, k7, fma(w45.x, k6w[6], fma(w23.y, k5w[6], fma(w23.x, k4w[6], fma(w01.y, k3w[6], fma(w01.x, k1w[6], yw[6]))))));
                st_if(a.phi[ee] + (g + phi_shift), ph, ok);
            }
        }
#elif defined(OC_VREGCONST)
        {
            const int row = r - 6;
            const bool ok = col_out && row >= y0 && row < y1;
            const int g = (row - a.row_base) * a.Nx + gx;
            double e = fma(hev[5], k7, fma(hev[4], k6w[6], fma(hev[3], k5w[6], fma(hev[2], k4w[6], fma(hev[1], k3w[6], hev[0] * k1w[6])))));
            double sc = fma(fmax(fabs(yw[6]), fabs(un[6])), a.rtol, a.atol);
            double qq = e * rcp_pos(sc);
            acc = ok ? fma(qq, qq, acc) : acc;
            st_if(a.ynew + g, un[6], ok);
            st_if(a.k7 + g, k7, ok);
            if (P2P) {  // rows within HY of a band edge also go to the neighbour's halo rows (NVLink peer stores)
                const bool up = ok && row < a.own0 + HY && a.peer_ynew[0] != nullptr;
                const bool dn = ok && row >= a.own1 - HY && a.peer_ynew[1] != nullptr;
                const int gu = (row - a.peer_row_base[0]) * a.Nx + gx, gd = (row - a.peer_row_base[1]) * a.Nx + gx;
                st_if(a.peer_ynew[0] + gu, un[6], up);
                st_if(a.peer_k7[0] + gu, k7, up);
                st_if(a.peer_ynew[1] + gd, un[6], dn);
                st_if(a.peer_k7[1] + gd, k7, dn);
            }
#pragma unroll
            for (int ee = 0; ee < NE; ee++) {
                double ph;
                if (ee < NV)
                    ph = fma(wv[ee][5], k7, fma(wv[ee][4], k6w[6], fma(wv[ee][3], k5w[6], fma(wv[ee][2], k4w[6],
                             fma(wv[ee][1], k3w[6], fma(wv[ee][0], k1w[6], yw[6]))))));
                else
                    ph = fma(a.w[ee][6], k7, fma(a.w[ee][5], k6w[6], fma(a.w[ee][4], k5w[6],
                             fma(a.w[ee][3], k4w[6], fma(a.w[ee][2], k3w[6], fma(a.w[ee][0], k1w[6], yw[6]))))));
                st_if(a.phi[ee] + (g + phi_shift), ph, ok);
            }
        }
#elif defined(OC_BRANCHLESS)
        {
            // the arithmetic is unconditional (halo threads compute values nobody reads) and only the stores and the
            // error accumulation are predicated: no divergent region inside the unrolled body
            const int row = r - 6;
            const bool ok = col_out && row >= y0 && row < y1;
            const int g = (row - a.row_base) * a.Nx + gx;
            double e = fma(a.he[6], k7, fma(a.he[5], k6w[6], fma(a.he[4], k5w[6], fma(a.he[3], k4w[6],
                           fma(a.he[2], k3w[6], a.he[0] * k1w[6])))));
            double sc = fma(fmax(fabs(yw[6]), fabs(un[6])), a.rtol, a.atol);
            double qq = e * rcp_pos(sc);
            acc = ok ? fma(qq, qq, acc) : acc;
            st_if(a.ynew + g, un[6], ok);
            st_if(a.k7 + g, k7, ok);
#pragma unroll
            for (int ee = 0; ee < NE; ee++) {
                double ph = fma(a.w[ee][6], k7, fma(a.w[ee][5], k6w[6], fma(a.w[ee][4], k5w[6],
                                fma(a.w[ee][3], k4w[6], fma(a.w[ee][2], k3w[6], fma(a.w[ee][0], k1w[6], yw[6]))))));
                st_if(a.phi[ee] + (g + phi_shift), ph, ok);
            }
        }
#else
        {
            const int row = r - 6;
            if (col_out && row >= y0 && row < y1) {
                const int g = (row - a.row_base) * a.Nx + gx;
                a.ynew[g] = un[6];
                a.k7[g] = k7;
                // rk.py:106,146-147: err = h * K.E ; scale = atol + max(|y|,|y_new|) * rtol
                double e = fma(a.he[6], k7, fma(a.he[5], k6w[6], fma(a.he[4], k5w[6], fma(a.he[3], k4w[6],
                               fma(a.he[2], k3w[6], a.he[0] * k1w[6])))));
                double sc = fma(fmax(fabs(yw[6]), fabs(un[6])), a.rtol, a.atol);
                double qq = e * rcp_pos(sc);
                acc = fma(qq, qq, acc);
#pragma unroll
                for (int ee = 0; ee < NE; ee++) {
                    // rk.py:723-737: y_old + h * Q.p with Q = K^T P, regrouped per stage
                    double ph = fma(a.w[ee][6], k7, fma(a.w[ee][5], k6w[6], fma(a.w[ee][4], k5w[6],
                                    fma(a.w[ee][3], k4w[6], fma(a.w[ee][2], k3w[6], fma(a.w[ee][0], k1w[6], yw[6]))))));
                    a.phi[ee][g + phi_shift] = ph;
                }
            }
        }
#endif
        __syncthreads();
        // ---- shift windows by one row (pure renaming inside the unrolled body)
#pragma unroll
        for (int l = 6; l > 0; l--) { yw[l] = yw[l - 1]; k1w[l] = k1w[l - 1]; cw[l] = cw[l - 1]; }
        k2w[4] = k2w[3]; k2w[3] = k2w[2]; k2w[2] = k2w[1];
        k3w[6] = k3w[5]; k3w[5] = k3w[4]; k3w[4] = k3w[3]; k3w[3] = k3w[2];
        k4w[6] = k4w[5]; k4w[5] = k4w[4]; k4w[4] = k4w[3];
        k5w[6] = k5w[5]; k5w[5] = k5w[4];
        k6w[6] = k6w[5];
        u2[2] = u2[1]; u2[1] = u2[0];
        u3[3] = u3[2]; u3[2] = u3[1];
        u4[4] = u4[3]; u4[3] = u4[2];
        u5[5] = u5[4]; u5[4] = u5[3];
        u6[6] = u6[5]; u6[5] = u6[4];
        un[7] = un[6]; un[6] = un[5];
      }
    }
    cp_async_wait<0>();
    // fixed-order CTA reduction of the error partial sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) sm.red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < BX / 32; i++) s += sm.red[i];
        a.partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
    if (a.ticket == nullptr) return;
    // ---- last CTA: fixed-order final reduction (see Args)
    __shared__ int is_last;
    if (tid == 0) {
        if (P2P) __threadfence_system();  // this CTA's peer stores are ordered before its ticket
        else __threadfence();
        const unsigned t = atomicAdd(a.ticket, 1u);
        is_last = (t == gridDim.x * gridDim.y - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double *rows = &sm.ex[0][0][0];
    const int ngx = gridDim.x, gy = gridDim.y, lane = tid & 31;
    for (int row = tid >> 5; row < gy; row += BX / 32) {
        // the 8 warps of rowgroup_sum_kernel's 256 threads, emulated by this warp: all loads first (independent, in
        // flight together), then the 8 shuffle trees, then the left-to-right sum of the 8 warp totals
        double v[8];
#pragma unroll
        for (int vw = 0; vw < 8; vw++) {
            v[vw] = 0.0;
            for (int i = vw * 32 + lane; i < ngx; i += 256) v[vw] += __ldcg(a.partial + (size_t)row * ngx + i);
        }
        double srow = 0.0;
#pragma unroll
        for (int vw = 0; vw < 8; vw++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[vw] += __shfl_down_sync(0xffffffffu, v[vw], o);
            srow += v[vw];  // meaningful on lane 0
        }
        if (lane == 0) rows[row] = srow;
    }
    __syncthreads();
    if (P2P) {
        // all-gather of the chunk-row sums through the ranks' inboxes (double buffered by the parity of seq)
        const int par = (int)(a.seq & 1ull);
        if (tid < a.nranks) {
            double *dst = a.peer_inbox[tid] + (size_t)(par * P2P_MAX_RANKS + a.my_rank) * P2P_INBOX_STRIDE;
            for (int r = 0; r < gy; r++) *reinterpret_cast<volatile double *>(dst + r) = rows[r];
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(dst + P2P_INBOX_STRIDE - 1) = a.seq;
            const volatile unsigned long long *flag = reinterpret_cast<const volatile unsigned long long *>(
                a.peer_inbox[a.my_rank] + (size_t)(par * P2P_MAX_RANKS + tid) * P2P_INBOX_STRIDE + P2P_INBOX_STRIDE - 1);
            // bounded wait (>= 8 s): a rank that died must not leave the others spinning in a kernel for ever
            long long it = 0;
            while (*flag != a.seq) {
                if (++it > 40000000ll) { is_last = 2; break; }
                __nanosleep(200);
            }
            __threadfence_system();
        }
        __syncthreads();
        if (tid == 0 && is_last == 2) {  // peer time-out: tell the host (sequence number with the top bit set)
            *a.ticket = 0u;
            *reinterpret_cast<volatile unsigned long long *>(a.result_seq) = a.seq | (1ull << 63);
        } else if (tid == 0) {
            double tot = 0.0;
            for (int q = 0; q < a.nranks; q++) {  // global chunk order: ranks top to bottom, chunk rows top to bottom
                const volatile double *src = a.peer_inbox[a.my_rank] + (size_t)(par * P2P_MAX_RANKS + q) * P2P_INBOX_STRIDE;
                for (int r = 0; r < gy; r++) tot += src[r];
            }
            *a.ticket = 0u;
            *reinterpret_cast<volatile double *>(a.result) = tot;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(a.result_seq) = a.seq;
        }
        return;
    }
    if (tid == 0) {
        double tot = 0.0;
        for (int r = 0; r < gy; r++) tot += rows[r];
        *a.ticket = 0u;
        *reinterpret_cast<volatile double *>(a.result) = tot;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(a.result_seq) = a.seq;
    }
}

}  // namespace fused
