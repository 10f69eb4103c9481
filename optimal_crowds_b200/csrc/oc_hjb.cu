// oc_hjb.cu -- HJB solve on sm_100a: RHS stencil (K1), RK45 stage kernels with the stage combination
// fused into the stencil load (K2, stage-wise formulation), dense output + velocity epilogue (K3) and
// the host-side replica of scipy's RK45 controller.
//
// Reference: optimals.py:124-206 (compute_optimal_velocity) + scipy 1.18.1 solve_ivp/RK45
// (rk.py:14-71,85-180,538-565,715-737; common.py:63-134; ivp.py:604-621,659-728; base.py:179-210).
//
// Data layout in HBM: every field is a dense C-order (Ny,Nx) FP64 array.  Workspace = y, y_new, K[0..6],
// coef (10 arrays).  coef = (V + g*m)/(mu*sigma^2), with NaN marking wall cells (V<0 -> k = 0, optimals.py:162).
//
// Error norm: per-tile partial sums -> per-tile-row sums (fixed order) -> sequential sum over tile rows on
// the host.  The order depends only on global tile indices, so a row-band decomposition whose bands are
// multiples of TY rows reproduces the single-GPU sum bit for bit.
#include <cmath>
#include <algorithm>
#include <vector>

#include "oc_common.h"
#include "oc_hjb_fused.cuh"
#include "oc_hjb_final.h"
#include "oc_rk45.h"
#include "oc_vels.h"

namespace {

#define RK_A rk45::A
#define RK_B rk45::B
#define RK_E rk45::E
#define RK_P rk45::P

constexpr int TX = 128, TY = 16, NTHREADS = 256;  // stage tile
constexpr int SW = TX + 2;                        // smem row pitch (doubles)

// y + h * sum_j a[j]*k[j]  (rk.py:63-64: dy = np.dot(K[:s].T, a[:s]) * h ; y + dy)
struct Comb {
    const double *y;
    const double *k[6];
    double a[6];
    double h;
};

template <int N>
__device__ __forceinline__ double comb_eval(const Comb &c, size_t idx) {
    if (N == 0) return c.y[idx];
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++) acc += __ldg(c.k[j] + idx) * c.a[j];
    return c.y[idx] + acc * c.h;
}

struct ErrArgs {        // MODE 1 only
    const double *k[6];  // K[0..5] (K[6] is the kernel's own output)
    double e[7];
    double h, rtol, atol;
};

__device__ __forceinline__ double block_sum(double v, double *red) {
    // fixed-order tree: warp shuffle, then warp 0 over the 8 warp sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NTHREADS / 32; i++) s += red[i];
    }
    return s;  // valid on thread 0
}

// MODE 0: kout = f(comb)                         (one RK stage, rk.py:62-64)
// MODE 1: ynew = comb; kout = f(ynew); partial[tile] = sum((h*K.E/scale)^2)   (rk.py:66-69,146-147)
template <int N, int MODE>
__global__ void __launch_bounds__(NTHREADS)
hjb_stage_kernel(Comb c, const double *__restrict__ coef, double *__restrict__ kout, double *__restrict__ ynew,
                 ErrArgs ea, double *__restrict__ partial, int Ny, int Nx, double diff_over_dxdy) {
    __shared__ double tile[(TY + 2) * SW];
    __shared__ double red[NTHREADS / 32];
    const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY;
    for (int idx = threadIdx.x; idx < (TY + 2) * SW; idx += NTHREADS) {
        int ly = idx / SW, lx = idx - ly * SW;
        int gy = ty0 + ly - 1, gx = tx0 + lx - 1;
        // mirror ghosts (optimals.py:149-152): row -1 := row 1, row Ny := row Ny-2; same for columns
        if (gy < 0) gy = 1;
        if (gy == Ny) gy = Ny - 2;
        if (gx < 0) gx = 1;
        if (gx == Nx) gx = Nx - 2;
        double v = 0.0;
        if (gy < Ny && gx < Nx) v = comb_eval<N>(c, (size_t)gy * Nx + gx);
        tile[idx] = v;
    }
    __syncthreads();
    double acc = 0.0;
    const int lx = threadIdx.x & (TX - 1);
    const int gx = tx0 + lx;
#pragma unroll 2
    for (int ly = threadIdx.x / TX; ly < TY; ly += NTHREADS / TX) {
        int gy = ty0 + ly;
        if (gy >= Ny || gx >= Nx) continue;
        size_t g = (size_t)gy * Nx + gx;
        const double *t = tile + (ly + 1) * SW + (lx + 1);
        double C = t[0];
        // optimals.py:154-160 (same summation order: up + down + left + right - 4C)
        double lap = t[-SW] + t[SW] + t[-1] + t[1] - 4.0 * C;
        double cf = coef[g];
        double r = diff_over_dxdy * lap - cf * C;
        if (cf != cf) r = 0.0;  // wall: optimals.py:162
        kout[g] = r;
        if (MODE == 1) {
            ynew[g] = C;
            double e = r * ea.e[6];
#pragma unroll
            for (int j = 5; j >= 0; j--)
                if (j != 1) e += __ldg(ea.k[j] + g) * ea.e[j];
            double sc = ea.atol + fmax(fabs(c.y[g]), fabs(C)) * ea.rtol;
            double q = (e * ea.h) / sc;
            acc += q * q;
        }
    }
    if (MODE == 1) {
        double s = block_sum(acc, red);
        if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

// sum of the tile partials of one tile row, in fixed order -> rowsum[tile_row]
__global__ void __launch_bounds__(NTHREADS) rowgroup_sum_kernel(const double *__restrict__ partial, int nbx,
                                                                double *__restrict__ rowsum) {
    __shared__ double red[NTHREADS / 32];
    double v = 0.0;
    for (int i = threadIdx.x; i < nbx; i += NTHREADS) v += partial[(size_t)blockIdx.x * nbx + i];
    double s = block_sum(v, red);
    if (threadIdx.x == 0) rowsum[blockIdx.x] = s;
}

// select_initial_step norms (common.py:109-127): mode 0: (y/sc)^2 and (f/sc)^2 ; mode 1: ((f1-f0)/sc)^2
__global__ void __launch_bounds__(NTHREADS)
init_norm_kernel(const double *__restrict__ y, const double *__restrict__ f0, const double *__restrict__ f1,
                 double rtol, double atol, int Ny, int Nx, double *__restrict__ partial, int mode) {
    __shared__ double red[NTHREADS / 32];
    const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY;
    const int gx = tx0 + (threadIdx.x & (TX - 1));
    double a0 = 0.0, a1 = 0.0;
    for (int ly = threadIdx.x / TX; ly < TY; ly += NTHREADS / TX) {
        int gy = ty0 + ly;
        if (gy >= Ny || gx >= Nx) continue;
        size_t g = (size_t)gy * Nx + gx;
        double sc = atol + fabs(y[g]) * rtol;
        if (mode == 0) {
            double q0 = y[g] / sc, q1 = f0[g] / sc;
            a0 += q0 * q0;
            a1 += q1 * q1;
        } else {
            double q = (f1[g] - f0[g]) / sc;
            a0 += q * q;
        }
    }
    size_t nb = (size_t)gridDim.x * gridDim.y, b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    double s0 = block_sum(a0, red);
    if (threadIdx.x == 0) partial[b] = s0;
    __syncthreads();
    if (mode == 0) {
        double s1 = block_sum(a1, red);
        if (threadIdx.x == 0) partial[nb + b] = s1;
    }
}

// coef = (V<0) ? NaN : (V + g*m)/(mu*sigma^2); y = 1 (optimals.py:83)
__global__ void prep_kernel(const double *__restrict__ V, const double *__restrict__ m, double g, double inv_den,
                            size_t n, double *__restrict__ coef, double *__restrict__ y) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = V[i];
    double mm = m ? m[i] : 0.0;
    coef[i] = (v < 0) ? __longlong_as_double(0x7ff8000000000000LL) : (v + g * mm) * inv_den;
    if (y) y[i] = 1.0;
}

// ---------------------------------------------------------------------------------------------------
// K3: dense output (rk.py:178-180,723-737) + vels (optimals.py:168-186), up to DMAX t_eval points per launch.
constexpr int DTX = 64, DTY = 8, DSW = DTX + 2, DMAX = 6;
constexpr int DCELLS = (DTY + 2) * DSW;                      // 660
constexpr int DPER = (DCELLS + NTHREADS - 1) / NTHREADS;     // 3

struct DenseArgs {
    const double *yold;
    const double *k[7];
    double P[7][4];
    double h;  // t - t_old (RkDenseOutput.h)
    int n_emit;
    double x[DMAX];       // (t_eval - t_old)/h
    double *phi[DMAX];    // (Ny,Nx) slice or NULL
    double *vx[DMAX];     // (Ny-2,Nx-2) slice or NULL
    double *vy[DMAX];
    double mu, lim, inv_two_dx, inv_two_dy;
};

__global__ void __launch_bounds__(NTHREADS) hjb_dense_kernel(DenseArgs a, int Ny, int Nx) {
    __shared__ double tile[DCELLS];
    const int tx0 = blockIdx.x * DTX, ty0 = blockIdx.y * DTY;
    double q[DPER][4], y0[DPER];
#pragma unroll
    for (int s = 0; s < DPER; s++) {
        int idx = threadIdx.x + s * NTHREADS;
        q[s][0] = q[s][1] = q[s][2] = q[s][3] = 0.0;
        y0[s] = 0.0;
        if (idx < DCELLS) {
            int ly = idx / DSW, lx = idx - ly * DSW;
            int gy = ty0 + ly - 1, gx = tx0 + lx - 1;
            if (gy >= 0 && gy < Ny && gx >= 0 && gx < Nx) {
                size_t g = (size_t)gy * Nx + gx;
                y0[s] = a.yold[g];
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    if (j == 1) continue;  // P[1,:] == 0
                    double kj = __ldg(a.k[j] + g);
#pragma unroll
                    for (int p = 0; p < 4; p++) q[s][p] += kj * a.P[j][p];
                }
            }
        }
    }
    for (int e = 0; e < a.n_emit; e++) {
        double x = a.x[e];
        double p1 = x, p2 = p1 * x, p3 = p2 * x, p4 = p3 * x;  // np.cumprod
        __syncthreads();
#pragma unroll
        for (int s = 0; s < DPER; s++) {
            int idx = threadIdx.x + s * NTHREADS;
            if (idx < DCELLS) {
                double d = q[s][0] * p1 + q[s][1] * p2 + q[s][2] * p3 + q[s][3] * p4;
                double ph = a.h * d + y0[s];
                int ly = idx / DSW, lx = idx - ly * DSW;
                int gy = ty0 + ly - 1, gx = tx0 + lx - 1;
                if (a.phi[e] && ly >= 1 && ly <= DTY && lx >= 1 && lx <= DTX && gy < Ny && gx < Nx)
                    a.phi[e][(size_t)gy * Nx + gx] = ph;
                tile[idx] = clamp_lim(ph, a.lim);
            }
        }
        __syncthreads();
        if (a.vx[e]) {
#pragma unroll
            for (int r = 0; r < (DTX * DTY) / NTHREADS; r++) {
                int cell = threadIdx.x + r * NTHREADS;
                int ly = cell / DTX, lx = cell - ly * DTX;
                int gy = ty0 + ly, gx = tx0 + lx;
                if (gy >= 1 && gy < Ny - 1 && gx >= 1 && gx < Nx - 1) {
                    const double *t = tile + (ly + 1) * DSW + (lx + 1);
                    double ox, oy;
                    vels_point(t[0], t[-1], t[1], t[-DSW], t[DSW], a.mu, a.lim, a.inv_two_dx, a.inv_two_dy, ox, oy);
                    size_t o = (size_t)(gy - 1) * (Nx - 2) + (gx - 1);
                    a.vx[e][o] = ox;
                    a.vy[e][o] = oy;
                }
            }
        }
    }
}

// standalone vels (optimals.py:168-186) for unit parity
__global__ void __launch_bounds__(NTHREADS)
vels_kernel(const double *__restrict__ phi, int Ny, int Nx, double mu, double lim, double inv_two_dx,
            double inv_two_dy, double *__restrict__ vx, double *__restrict__ vy) {
    int gx = blockIdx.x * 64 + (threadIdx.x & 63) + 1, gy = blockIdx.y * 4 + (threadIdx.x >> 6) + 1;
    if (gy >= Ny - 1 || gx >= Nx - 1) return;
    size_t g = (size_t)gy * Nx + gx;
    double ox, oy;
    vels_point(clamp_lim(phi[g], lim), clamp_lim(phi[g - 1], lim), clamp_lim(phi[g + 1], lim),
               clamp_lim(phi[g - Nx], lim), clamp_lim(phi[g + Nx], lim), mu, lim, inv_two_dx, inv_two_dy, ox, oy);
    size_t o = (size_t)(gy - 1) * (Nx - 2) + (gx - 1);
    vx[o] = ox;
    vy[o] = oy;
}

// ---------------------------------------------------------------------------------------------------
struct Solver {
    oc_ctx *ctx;
    cudaStream_t st;
    int Ny, Nx, nbx, nby;
    size_t n;
    double *y, *ynew, *K[7], *coef;
    double *partial, *rowsum_d, *rowsum_h;
    double diff_over_dxdy;
    int launches = 0;
    bool profile = false;
    struct Rec { int cls; double bytes; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    size_t pool_used = 0;
    cudaEvent_t get_event() {
        if (pool_used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[pool_used++];
    }
    void begin(int cls, double bytes) {
        if (!profile) return;
        Rec r{cls, bytes, get_event(), get_event()};
        cudaEventRecord(r.a, st);
        recs.push_back(r);
    }
    void end() {
        if (!profile) return;
        cudaEventRecord(recs.back().b, st);
    }
    void collect(oc_hjb_stats *out) {
        for (auto &r : recs) {
            float ms = 0;
            cudaEventElapsedTime(&ms, r.a, r.b);
            out->cls_launches[r.cls]++;
            out->cls_ms[r.cls] += ms;
            out->cls_bytes[r.cls] += r.bytes;
        }
        for (auto e : pool) cudaEventDestroy(e);
        pool.clear(); recs.clear();
    }

    template <int N, int MODE>
    void launch_stage(const Comb &c, double *kout, double *yn, const ErrArgs &ea) {
        dim3 grid(nbx, nby);
        // algorithmic words per cell: y + N k's + coef in, k out (+ y_new out, K[0..5] re-used for the error)
        begin(0, 8.0 * (double)n * (MODE == 1 ? 9 : N + 3));
        hjb_stage_kernel<N, MODE><<<grid, NTHREADS, 0, st>>>(c, coef, kout, yn, ea, partial, Ny, Nx, diff_over_dxdy);
        end();
        launches++;
    }
    void stage(int N, const Comb &c, double *kout) {
        ErrArgs ea{};
        switch (N) {
            case 0: launch_stage<0, 0>(c, kout, nullptr, ea); break;
            case 1: launch_stage<1, 0>(c, kout, nullptr, ea); break;
            case 2: launch_stage<2, 0>(c, kout, nullptr, ea); break;
            case 3: launch_stage<3, 0>(c, kout, nullptr, ea); break;
            case 4: launch_stage<4, 0>(c, kout, nullptr, ea); break;
            case 5: launch_stage<5, 0>(c, kout, nullptr, ea); break;
        }
    }
    // ---- stage-fused step (oc_hjb_fused.cuh)
    int fused_rc = 0, fused_gx = 0, fused_gy = 0;
    void fused_plan(int n_sm, int forced_rc = 0) {
        fused_gx = (Nx + fused::VX - 1) / fused::VX;
        fused_rc = forced_rc > 0 ? forced_rc : fused::plan_chunk_rows(Nx, Ny, n_sm, 1, 0);
        fused_gy = (Ny + fused_rc - 1) / fused_rc;
    }
    fused::MapCache maps;  // TMA descriptors of this solve's arrays
    int launch_fused(int ne, fused::Args &a) {
        fused::set_tensor_maps(a, Ny, &maps);
        dim3 grid(fused_gx, fused_gy);
        begin(0, 8.0 * (double)n * (5 + ne));  // y, f, coef in; y_new, f_new, NE phi slices out
        OC_CUDA(fused::launch<false>(ne, a, grid, st));
        end();
        launches++;
        return OC_OK;
    }
    // sum over tile rows of per-tile partials located at partial+off; host-visible after sync
    int reduce_to_host(size_t off, double *out, int gx = -1, int gy = -1) {
        if (gx < 0) { gx = nbx; gy = nby; }
        begin(2, 8.0 * ((double)gx * gy + gy));
        rowgroup_sum_kernel<<<gy, NTHREADS, 0, st>>>(partial + off, gx, rowsum_d);
        end();
        launches++;
        OC_CUDA(cudaMemcpyAsync(rowsum_h, rowsum_d, sizeof(double) * gy, cudaMemcpyDeviceToHost, st));
        OC_CUDA(cudaStreamSynchronize(st));
        double s = 0.0;
        for (int i = 0; i < gy; i++) s += rowsum_h[i];  // fixed global order
        *out = s;
        return OC_OK;
    }
};

int ensure_ws(oc_ctx *ctx, size_t n, int nbx, int nby) {
    size_t need = 10 * n * sizeof(double);
    if (ctx->hjb_ws_bytes < need) {
        if (ctx->hjb_ws) cudaFree(ctx->hjb_ws);
        ctx->hjb_ws = nullptr;
        ctx->hjb_ws_bytes = 0;
        if (cudaMalloc(&ctx->hjb_ws, need) != cudaSuccess) {
            cudaGetLastError();
            oc::set_error("cannot allocate %zu bytes of HJB workspace", need);
            return OC_ERR_NOMEM;
        }
        ctx->hjb_ws_bytes = need;
    }
    // large enough for the stage tiles and for any fused-step plan (>= 32-row chunks, 240-column tiles)
    size_t np = (size_t)2 * nbx * nby + nby + 64;
    {
        size_t fgx = (size_t)(nbx * TX + fused::VX - 1) / fused::VX + 1, fgy = (size_t)(nby * TY + 15) / 16 + 1;
        np = std::max(np, 2 * fgx * fgy + fgy + 64);
    }
    if (ctx->hjb_partial_n < np) {
        if (ctx->hjb_partial) cudaFree(ctx->hjb_partial);
        OC_CUDA(cudaMalloc(&ctx->hjb_partial, np * sizeof(double)));
        ctx->hjb_partial_n = np;
    }
    size_t hp = (size_t)std::max(nby, (nby * TY + 15) / 16 + 1) + 16;
    if (ctx->h_pinned_n < hp) {
        if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
        OC_CUDA(cudaMallocHost(&ctx->h_pinned, hp * sizeof(double)));
        ctx->h_pinned_n = hp;
    }
    return OC_OK;
}

int ensure_final_reduce(oc_ctx *ctx, int slots) { return ocfinal::ensure(ctx, slots); }
void final_reduce_args(oc_ctx *ctx, int slot, fused::Args &fa) { ocfinal::set_args(ctx, slot, ++ctx->fr_seq, fa); }
inline bool final_ready(oc_ctx *ctx, int slot, unsigned long long seq, double *out) { return ocfinal::ready(ctx, slot, seq, out); }
int wait_final(oc_ctx *ctx, int slot, unsigned long long seq, cudaStream_t st, double *out) {
    return ocfinal::wait(ctx, slot, seq, st, out);
}

}  // namespace

extern "C" int oc_hjb_rhs(oc_ctx *ctx, const double *d_phi, const double *d_V, const double *d_m,
                          const oc_hjb_params *prm, double *d_out, void *stream) {
    OC_ARG(ctx && d_phi && d_V && prm && d_out, "NULL argument");
    OC_CUDA(cudaSetDevice(ctx->device));
    Solver s{};
    s.ctx = ctx; s.st = (cudaStream_t)stream; s.Ny = ctx->Ny; s.Nx = ctx->Nx;
    s.n = (size_t)s.Ny * s.Nx;
    s.nbx = (s.Nx + TX - 1) / TX; s.nby = (s.Ny + TY - 1) / TY;
    int rc = ensure_ws(ctx, s.n, s.nbx, s.nby);
    if (rc) return rc;
    s.coef = ctx->hjb_ws;
    s.partial = ctx->hjb_partial;
    double s2 = prm->sigma * prm->sigma;
    s.diff_over_dxdy = (-0.5 * s2) / (ctx->dx * ctx->dy);
    prep_kernel<<<(unsigned)((s.n + 255) / 256), 256, 0, s.st>>>(d_V, d_m, prm->g, 1.0 / (prm->mu * s2), s.n, s.coef, nullptr);
    Comb c{};
    c.y = d_phi;
    s.stage(0, c, d_out);
    oc::count_launch(2);
    OC_CUDA(cudaGetLastError());
    return OC_OK;
}

// rows per CTA chunk the stage-fused kernel would choose for `copies` independent solves of this context's grid running
// side by side (ensembles).  A caller that wants member results independent of the batch size fixes prm->chunk_rows
// to this value for the stand-alone solve too (the chunking defines the order of the error-norm sum).
extern "C" int oc_hjb_plan_chunk_rows(oc_ctx *ctx, int copies) {
    if (!ctx || copies < 1) return OC_ERR_ARG;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
    return fused::plan_chunk_rows(ctx->Nx, ctx->Ny, n_sm, copies, 0);
}

extern "C" int oc_hjb_vels(oc_ctx *ctx, const double *d_phi, const oc_hjb_params *prm, double *d_vx, double *d_vy,
                           void *stream) {
    OC_ARG(ctx && d_phi && prm && d_vx && d_vy, "NULL argument");
    OC_CUDA(cudaSetDevice(ctx->device));
    dim3 grid((ctx->Nx - 2 + 63) / 64, (ctx->Ny - 2 + 3) / 4);
    vels_kernel<<<grid, NTHREADS, 0, (cudaStream_t)stream>>>(d_phi, ctx->Ny, ctx->Nx, prm->mu, prm->lim,
                                                             1.0 / (2 * ctx->dx), 1.0 / (2 * ctx->dy), d_vx, d_vy);
    oc::count_launch();
    OC_CUDA(cudaGetLastError());
    return OC_OK;
}

extern "C" int oc_hjb_solve(oc_ctx *ctx, const double *d_V, const double *d_m, const oc_hjb_params *prm, double T,
                            const double *t_eval, int nt, double *d_phi, double *d_vx, double *d_vy,
                            oc_hjb_stats *stats, double *trace_h, double *trace_err, int trace_cap, int *trace_n,
                            void *stream) {
    OC_ARG(ctx && d_V && prm && stats, "NULL argument");
    OC_ARG(nt >= 1 && t_eval, "t_eval required");
    OC_CUDA(cudaSetDevice(ctx->device));
    Solver s{};
    s.ctx = ctx; s.st = (cudaStream_t)stream; s.Ny = ctx->Ny; s.Nx = ctx->Nx;
    const size_t n = s.n = (size_t)s.Ny * s.Nx;
    s.nbx = (s.Nx + TX - 1) / TX; s.nby = (s.Ny + TY - 1) / TY;
    int rc = ensure_ws(ctx, n, s.nbx, s.nby);
    if (rc) return rc;
    double *ws = ctx->hjb_ws;
    s.coef = ws; s.y = ws + n; s.ynew = ws + 2 * n;
    for (int j = 0; j < 7; j++) s.K[j] = ws + (3 + j) * n;
    s.partial = ctx->hjb_partial;
    s.rowsum_d = ctx->hjb_partial + (size_t)2 * s.nbx * s.nby;
    s.rowsum_h = ctx->h_pinned;
    const double s2 = prm->sigma * prm->sigma;
    s.diff_over_dxdy = (-0.5 * s2) / (ctx->dx * ctx->dy);
    const double rtol = prm->rtol, atol = prm->atol;
    const double sqrt_n = std::sqrt((double)n);
    memset(stats, 0, sizeof(*stats));
    s.profile = prm->profile != 0;
    int ntr = 0;
    OC_CUDA(cudaEventRecord(ctx->ev0, s.st));

    prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s.st>>>(d_V, d_m, prm->g, 1.0 / (prm->mu * s2), n, s.coef, s.y);
    s.launches++;
    const double t_bound = 0.0;
    const double direction = (t_bound != T) ? (t_bound > T ? 1.0 : -1.0) : 1.0;
    double t = T;
    // rk.py:94: f = fun(t0, y0)
    {
        Comb c{};
        c.y = s.y;
        s.stage(0, c, s.K[0]);
        stats->nfev++;
    }
    // common.py:109-134 select_initial_step
    double h_abs;
    {
        double interval = std::fabs(t_bound - T);
        if (interval == 0.0) h_abs = 0.0;
        else {
            dim3 grid(s.nbx, s.nby);
            init_norm_kernel<<<grid, NTHREADS, 0, s.st>>>(s.y, s.K[0], nullptr, rtol, atol, s.Ny, s.Nx, s.partial, 0);
            s.launches++;
            double s0, s1, sq2;
            if ((rc = s.reduce_to_host(0, &s0))) return rc;
            if ((rc = s.reduce_to_host((size_t)s.nbx * s.nby, &s1))) return rc;
            double d0 = std::sqrt(s0) / sqrt_n, d1 = std::sqrt(s1) / sqrt_n;
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            h0 = std::min(h0, interval);
            Comb c{};
            c.y = s.y; c.k[0] = s.K[0]; c.a[0] = 1.0; c.h = h0 * direction;  // y1 = y0 + h0*direction*f0
            s.stage(1, c, s.K[1]);
            stats->nfev++;
            init_norm_kernel<<<grid, NTHREADS, 0, s.st>>>(s.y, s.K[0], s.K[1], rtol, atol, s.Ny, s.Nx, s.partial, 1);
            s.launches++;
            if ((rc = s.reduce_to_host(0, &sq2))) return rc;
            double d2 = std::sqrt(sq2) / sqrt_n / h0;
            double h1;
            if (d1 <= 1e-15 && d2 <= 1e-15) h1 = std::max(1e-6, h0 * 1e-3);
            else h1 = std::pow(0.01 / std::max(d1, d2), 1.0 / 5.0);
            h_abs = std::min(std::min(100 * h0, h1), interval);
        }
    }
    stats->h0 = h_abs;

    // stage-fused path (prm->fused): needs room for NE_MAX phi slices when only velocities are requested
    // the fused kernel mirrors up to 6 rows / 8 columns across the boundary: tiny grids take the stage-wise path
    bool use_fused = prm->fused != 0 && s.Ny > fused::HY + 1 && s.Nx > fused::HX + 1, fused_attempt = false;
    double *phi_scratch = nullptr;
    if (use_fused) {
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
        s.fused_plan(n_sm, prm->chunk_rows);
        size_t np_need = (size_t)2 * s.fused_gx * s.fused_gy + s.fused_gy;
        if (ctx->hjb_partial_n < np_need || ctx->h_pinned_n < (size_t)s.fused_gy + 16) {
            oc::set_error("internal: fused partial buffers too small");
            return OC_ERR_ARG;
        }
        if (!d_phi && d_vx && d_vy) {
            size_t need = (size_t)fused::NE_MAX * n * sizeof(double);
            if (ctx->hjb_scratch_bytes < need) {
                if (ctx->hjb_scratch) cudaFree(ctx->hjb_scratch);
                ctx->hjb_scratch = nullptr; ctx->hjb_scratch_bytes = 0;
                if (cudaMalloc(&ctx->hjb_scratch, need) != cudaSuccess) {
                    cudaGetLastError();
                    oc::set_error("cannot allocate %zu bytes of phi scratch", need);
                    return OC_ERR_NOMEM;
                }
                ctx->hjb_scratch_bytes = need;
            }
            phi_scratch = ctx->hjb_scratch;
        }
    }
    int t_eval_i = nt;  // ivp.py:617-621
    int n_out = 0, status = 1;
    const double error_exponent = -1.0 / 5.0;
    const size_t n_int = (size_t)(s.Ny - 2) * (s.Nx - 2);
    const bool final_in_kernel = use_fused && s.fused_gy <= fused::MAX_FINAL_ROWS;
    if (final_in_kernel && (rc = ensure_final_reduce(ctx, 1))) return rc;

    while (status == 1) {
        if (t == t_bound) { status = 0; break; }  // base.py:193-198
        double min_step = 10 * std::fabs(std::nextafter(t, direction * INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        bool accepted = false, rejected = false;
        double h = 0, t_new = 0;
        while (!accepted) {
            if (h_abs < min_step) { status = -1; break; }
            h = h_abs * direction;
            if (prm->forced_h && ntr < prm->n_forced_h) h = prm->forced_h[ntr];  // teacher-forced (tests)
            t_new = t + h;
            if (direction * (t_new - t_bound) > 0) t_new = t_bound;
            h = t_new - t;
            h_abs = std::fabs(h);
            // t_eval samples this attempt would emit if accepted (ivp.py:715-728): ascending indices
            // [ia_lo, t_eval_i) with t_eval >= t_new
            int ia_lo;
            {
                int lo = 0, hi = nt;
                while (lo < hi) {
                    int mid = (lo + hi) / 2;
                    if (t_eval[nt - 1 - mid] < t_new) lo = mid + 1; else hi = mid;
                }
                ia_lo = std::min(lo, t_eval_i);
            }
            const int n_emit = t_eval_i - ia_lo;
            const bool want_out = d_phi || (d_vx && d_vy);
            fused_attempt = use_fused && (!want_out || n_emit <= fused::NE_MAX);
            double se;
            if (fused_attempt) {
                // one launch: 6 RHS evaluations, y_new, f_new, error partial sums and the (speculative) dense
                // output samples; a rejected attempt's samples are overwritten by the step that finally covers them
                fused::Args fa{};
                fa.y = s.y; fa.k1 = s.K[0]; fa.coef = s.coef; fa.ynew = s.ynew; fa.k7 = s.K[6]; fa.partial = s.partial;
                fa.ha21 = h * RK_A[1][0];
                for (int j = 0; j < 2; j++) fa.ha3[j] = h * RK_A[2][j];
                for (int j = 0; j < 3; j++) fa.ha4[j] = h * RK_A[3][j];
                for (int j = 0; j < 4; j++) fa.ha5[j] = h * RK_A[4][j];
                for (int j = 0; j < 5; j++) fa.ha6[j] = h * RK_A[5][j];
                for (int j = 0; j < 6; j++) fa.hb[j] = h * RK_B[j];
                for (int j = 0; j < 7; j++) fa.he[j] = h * RK_E[j];
                fa.A = s.diff_over_dxdy; fa.rtol = rtol; fa.atol = atol;
                fa.Ny = s.Ny; fa.Nx = s.Nx; fa.RC = s.fused_rc;
                fa.row_base = 0; fa.own0 = 0; fa.own1 = s.Ny; fa.phi_row_base = 0;
                int ne = 0;
                if (want_out) {
                    for (int ia = t_eval_i - 1; ia >= ia_lo; ia--, ne++) {
                        const int kd = nt - 1 - ia;
                        const double x = (t_eval[kd] - t) / h;  // RkDenseOutput: (t - t_old)/h
                        double pw[4] = {x, x * x, x * x * x, x * x * x * x};
                        for (int j = 0; j < 7; j++) {
                            double acc = 0.0;
                            for (int q = 0; q < 4; q++) acc += RK_P[j][q] * pw[q];
                            fa.w[ne][j] = h * acc;
                        }
                        fa.phi[ne] = d_phi ? d_phi + (size_t)kd * n : phi_scratch + (size_t)ne * n;
                    }
                }
                if (final_in_kernel) final_reduce_args(ctx, 0, fa);
                if ((rc = s.launch_fused(ne, fa))) return rc;
                stats->nfev += 6;
                if (final_in_kernel) {
                    if ((rc = wait_final(ctx, 0, fa.seq, s.st, &se))) return rc;
                } else if ((rc = s.reduce_to_host(0, &se, s.fused_gx, s.fused_gy))) return rc;
            } else {
            // rk_step (rk.py:61-69): K[0] = f already in place
            for (int sgi = 1; sgi < 6; sgi++) {
                Comb c{};
                c.y = s.y; c.h = h;
                for (int j = 0; j < sgi; j++) { c.k[j] = s.K[j]; c.a[j] = RK_A[sgi][j]; }
                s.stage(sgi, c, s.K[sgi]);
            }
            {
                // y_new = y + h*dot(K[:-1].T,B) with B[1]=0 dropped; K[6] = f(y_new); error partials
                Comb c{};
                c.y = s.y; c.h = h;
                int m = 0;
                for (int j = 0; j < 6; j++)
                    if (j != 1) { c.k[m] = s.K[j]; c.a[m] = RK_B[j]; m++; }
                ErrArgs ea{};
                for (int j = 0; j < 6; j++) ea.k[j] = s.K[j];
                for (int j = 0; j < 7; j++) ea.e[j] = RK_E[j];
                ea.h = h; ea.rtol = rtol; ea.atol = atol;
                s.launch_stage<5, 1>(c, s.K[6], s.ynew, ea);
            }
            stats->nfev += 6;
            if ((rc = s.reduce_to_host(0, &se))) return rc;
            }
            double error_norm = std::sqrt(se) / sqrt_n;  // rk.py:147, common.py:63-65
            if (trace_h && ntr < trace_cap) { trace_h[ntr] = h; trace_err[ntr] = error_norm; }
            ntr++;
            if (error_norm < 1) {  // rk.py:149-161
                double factor = (error_norm == 0) ? 10.0 : std::min(10.0, 0.9 * std::pow(error_norm, error_exponent));
                if (rejected) factor = std::min(1.0, factor);
                h_abs *= factor;
                accepted = true;
                stats->n_accepted++;
            } else {  // rk.py:162-165
                h_abs *= std::max(0.2, 0.9 * std::pow(error_norm, error_exponent));
                rejected = true;
                stats->n_rejected++;
            }
        }
        if (status == -1) break;
        const double t_old = t;
        if (direction * (t_new - t_bound) >= 0) status = 0;  // base.py:205-208

        // ivp.py:715-728, direction < 0: first ascending index with t_eval_asc >= t_new
        int lo = 0, hi = nt;
        while (lo < hi) {
            int mid = (lo + hi) / 2;
            if (t_eval[nt - 1 - mid] < t_new) lo = mid + 1; else hi = mid;
        }
        const int t_eval_i_new = lo;
        if (fused_attempt) {
            // samples were written by the step kernel; only the gradient-to-velocity conversion is left
            int ne = 0;
            for (int ia = t_eval_i - 1; ia >= t_eval_i_new; ia--, ne++) {
                const int kd = nt - 1 - ia;
                n_out++;
                if (d_vx && d_vy && kd >= 1) {
                    const double *ph = d_phi ? d_phi + (size_t)kd * n : phi_scratch + (size_t)ne * n;
                    const int sl = nt - 1 - kd;
                    dim3 grid((s.Nx - 2 + 63) / 64, (s.Ny - 2 + 3) / 4);
                    s.begin(1, 8.0 * (double)n * 3);
                    vels_kernel<<<grid, NTHREADS, 0, s.st>>>(ph, s.Ny, s.Nx, prm->mu, prm->lim, 1.0 / (2 * ctx->dx), 1.0 / (2 * ctx->dy),
                                                             d_vx + (size_t)sl * n_int, d_vy + (size_t)sl * n_int);
                    s.end();
                    s.launches++;
                }
            }
            t_eval_i = t_eval_i_new;
        } else if (t_eval_i_new < t_eval_i) {
            const double hh = t_new - t_old;
            int ia = t_eval_i - 1;
            while (ia >= t_eval_i_new) {
                DenseArgs a{};
                a.yold = s.y;
                for (int j = 0; j < 7; j++) { a.k[j] = s.K[j]; for (int p = 0; p < 4; p++) a.P[j][p] = RK_P[j][p]; }
                a.h = hh; a.mu = prm->mu; a.lim = prm->lim; a.inv_two_dx = 1.0 / (2 * ctx->dx); a.inv_two_dy = 1.0 / (2 * ctx->dy);
                int e = 0;
                bool any = false;
                for (; e < DMAX && ia >= t_eval_i_new; ia--) {
                    int kd = nt - 1 - ia;  // column of sol.y
                    a.x[e] = (t_eval[kd] - t_old) / hh;
                    a.phi[e] = d_phi ? d_phi + (size_t)kd * n : nullptr;
                    int sl = nt - 1 - kd;  // vx_opt slice (optimals.py:200-204); column 0 (phi_T) has none
                    bool has_v = d_vx && d_vy && kd >= 1;
                    a.vx[e] = has_v ? d_vx + (size_t)sl * n_int : nullptr;
                    a.vy[e] = has_v ? d_vy + (size_t)sl * n_int : nullptr;
                    if (a.phi[e] || a.vx[e]) { any = true; e++; }
                    n_out++;
                }
                a.n_emit = e;
                if (any) {
                    dim3 grid((s.Nx + DTX - 1) / DTX, (s.Ny + DTY - 1) / DTY);
                    double words = 7.0;  // y_old + K[0], K[2..6]
                    for (int q = 0; q < e; q++) words += (a.phi[q] ? 1.0 : 0.0) + (a.vx[q] ? 2.0 : 0.0);
                    s.begin(1, 8.0 * (double)n * words);
                    hjb_dense_kernel<<<grid, NTHREADS, 0, s.st>>>(a, s.Ny, s.Nx);
                    s.end();
                    s.launches++;
                }
            }
            t_eval_i = t_eval_i_new;
        }
        // accept (rk.py:167-174): y <- y_new, f <- f_new (FSAL: pointer swaps)
        t = t_new;
        std::swap(s.y, s.ynew);
        std::swap(s.K[0], s.K[6]);
    }
    OC_CUDA(cudaEventRecord(ctx->ev1, s.st));
    OC_CUDA(cudaStreamSynchronize(s.st));
    OC_CUDA(cudaGetLastError());
    float ms = 0;
    OC_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    stats->gpu_ms = ms;
    s.collect(stats);
    stats->status = status;
    stats->n_out = n_out;
    stats->launches = s.launches;
    oc::count_launch(s.launches);
    if (trace_n) *trace_n = ntr;
    if (status == -1) {
        oc::set_error("RK45: required step size is less than spacing between numbers (t=%g)", t);
        return OC_ERR_STEP_TOO_SMALL;
    }
    return OC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Batched solve for ensembles of independent rooms on one grid shape (BASELINE configs[4]).
// Every room keeps its own RK45 controller (t, h, accept/reject, t_eval cursor: the reference integrates each
// room separately, optimals.py:193-196) and its own CUDA stream; a room's step attempt is ONE launch of the
// stage-fused kernel followed by the fixed-order reduction, exactly the launches oc_hjb_solve issues, so every
// room's result is bit-identical to solving it alone (with the same prm->chunk_rows).  The host serves the rooms
// round-robin: while it reads room b's error norm and enqueues b's next attempt, the other rooms' kernels keep
// the GPU busy -- small grids (512^2: ~10 us of device work per attempt) are launch-latency bound one at a time.
namespace {

struct BatchRoom {
    Solver s;
    enum Phase { START, INIT0, INIT1, STEP, DONE } phase = START;
    const double *V = nullptr, *m = nullptr;
    double *phi = nullptr, *vx = nullptr, *vy = nullptr;
    double *rs_d = nullptr, *rs_h = nullptr;  // row-group sums (device / pinned host), 3 regions of `rs_stride`
    int rs_stride = 0;
    cudaEvent_t ev = nullptr;
    unsigned long long wait_seq = 0;     // sequence number of the step attempt in flight (in-kernel final reduction)
    // controller state (rk.py:111-176, base.py:179-210, ivp.py:659-728)
    double t = 0, h_abs = 0, h = 0, t_new = 0, min_step = 0, h0 = 0, d1 = 0;
    bool rejected = false;
    int status = 1, t_eval_i = 0, n_out = 0, ia_lo = 0;
    double step_sum = 0.0;               // error sum delivered by the last CTA of the attempt
    fused::Args fa{};                    // the attempt waiting in a bucket for its batched launch
    oc_hjb_stats st{};
};

}  // namespace

extern "C" int oc_hjb_solve_batch(oc_ctx *ctx, int n_rooms, const double *const *d_V, const double *const *d_m,
                                  const oc_hjb_params *prm, double T, const double *t_eval, int nt,
                                  double *const *d_phi, double *const *d_vx, double *const *d_vy,
                                  oc_hjb_stats *stats, void *stream) {
    OC_ARG(ctx && d_V && prm && stats && n_rooms >= 1, "NULL argument");
    OC_ARG(nt >= 1 && t_eval, "t_eval required");
    OC_ARG(d_phi || (d_vx && d_vy), "an output (d_phi or d_vx,d_vy) is required");
    const int Ny = ctx->Ny, Nx = ctx->Nx;
    OC_ARG(Ny > fused::HY + 1 && Nx > fused::HX + 1, "grid too small for the stage-fused kernel");
    OC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t main_st = (cudaStream_t)stream;
    const size_t n = (size_t)Ny * Nx, n_int = (size_t)(Ny - 2) * (Nx - 2);
    const int nbx = (Nx + TX - 1) / TX, nby = (Ny + TY - 1) / TY;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
    const int fgx = (Nx + fused::VX - 1) / fused::VX;
    const int frc = prm->chunk_rows > 0 ? prm->chunk_rows : fused::plan_chunk_rows(Nx, Ny, n_sm, 1, 0);
    const int fgy = (Ny + frc - 1) / frc;
    // per room: coef, y, y_new, f, f_new (+ NE_MAX phi slices when only velocities are kept) ; partial sums
    const bool need_scratch = !d_phi;
    // (even numbers of doubles: every room's arrays stay 16-byte aligned for the TMA descriptors)
    const size_t np = ((std::max((size_t)2 * nbx * nby, (size_t)fgx * fgy) + 8) + 1) & ~(size_t)1;
    const int rs_stride = ((std::max(nby, fgy) + 8) + 1) & ~1;
    const size_t per_room = (5 + (need_scratch ? fused::NE_MAX : 0)) * n + np + 3 * (size_t)rs_stride;
    const size_t need = per_room * n_rooms * sizeof(double);
    if (ctx->batch_ws_bytes < need) {
        if (ctx->batch_ws) cudaFree(ctx->batch_ws);
        ctx->batch_ws = nullptr; ctx->batch_ws_bytes = 0;
        if (cudaMalloc(&ctx->batch_ws, need) != cudaSuccess) {
            cudaGetLastError();
            oc::set_error("cannot allocate %zu bytes of batched HJB workspace (%d rooms)", need, n_rooms);
            return OC_ERR_NOMEM;
        }
        ctx->batch_ws_bytes = need;
    }
    const size_t pin_need = (size_t)3 * rs_stride * n_rooms;
    if (ctx->batch_pinned_n < pin_need) {
        if (ctx->batch_pinned) cudaFreeHost(ctx->batch_pinned);
        ctx->batch_pinned = nullptr; ctx->batch_pinned_n = 0;
        OC_CUDA(cudaMallocHost(&ctx->batch_pinned, pin_need * sizeof(double)));
        ctx->batch_pinned_n = pin_need;
    }
    constexpr int MAX_STREAMS = 32;
    const int n_streams = std::min(n_rooms, MAX_STREAMS);
    while ((int)ctx->batch_streams.size() < n_streams) {
        cudaStream_t s;
        OC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        ctx->batch_streams.push_back(s);
    }
    while ((int)ctx->batch_events.size() < n_rooms + 1) {
        cudaEvent_t e;
        OC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->batch_events.push_back(e);
    }
    // inputs produced on the caller's stream are visible to the room streams
    OC_CUDA(cudaEventRecord(ctx->ev0, main_st));
    cudaEvent_t ev_in = ctx->batch_events[n_rooms];
    OC_CUDA(cudaEventRecord(ev_in, main_st));
    for (int q = 0; q < n_streams; q++) OC_CUDA(cudaStreamWaitEvent(ctx->batch_streams[q], ev_in, 0));

    const double s2 = prm->sigma * prm->sigma;
    const double rtol = prm->rtol, atol = prm->atol, sqrt_n = std::sqrt((double)n);
    const double t_bound = 0.0, direction = (t_bound != T) ? (t_bound > T ? 1.0 : -1.0) : 1.0;
    const double error_exponent = -1.0 / 5.0;
    const bool want_v = d_vx && d_vy;
    const bool final_in_kernel = fgy <= fused::MAX_FINAL_ROWS;
    if (final_in_kernel) {
        int frc = ensure_final_reduce(ctx, n_rooms);
        if (frc) return frc;
    }
    std::vector<BatchRoom> rooms(n_rooms);
    for (int b = 0; b < n_rooms; b++) {
        BatchRoom &r = rooms[b];
        OC_ARG(d_V[b] && (!d_phi || d_phi[b]) && (!want_v || (d_vx[b] && d_vy[b])), "NULL room pointer");
        double *ws = ctx->batch_ws + per_room * b;
        Solver &s = r.s;
        s.ctx = ctx; s.st = ctx->batch_streams[b % n_streams]; s.Ny = Ny; s.Nx = Nx; s.n = n; s.nbx = nbx; s.nby = nby;
        s.coef = ws; s.y = ws + n; s.ynew = ws + 2 * n; s.K[0] = ws + 3 * n; s.K[6] = ws + 4 * n;
        s.K[1] = s.K[6];  // f(y0 + h0 f0) of select_initial_step lives in the (still unused) f_new array
        double *scratch = need_scratch ? ws + 5 * n : nullptr;
        s.partial = ws + (5 + (need_scratch ? fused::NE_MAX : 0)) * n;
        r.rs_d = s.partial + np;
        r.rs_h = ctx->batch_pinned + (size_t)3 * rs_stride * b;
        r.rs_stride = rs_stride;
        s.diff_over_dxdy = (-0.5 * s2) / (ctx->dx * ctx->dy);
        s.fused_gx = fgx; s.fused_rc = frc; s.fused_gy = fgy;
        r.V = d_V[b]; r.m = d_m ? d_m[b] : nullptr;
        r.phi = d_phi ? d_phi[b] : scratch;
        r.vx = want_v ? d_vx[b] : nullptr; r.vy = want_v ? d_vy[b] : nullptr;
        r.ev = ctx->batch_events[b];
        r.t = T; r.t_eval_i = nt;
    }
    // room-local helpers ------------------------------------------------------------------------------
    auto reduce_async = [&](BatchRoom &r, size_t off, int gx, int gy, int slot) {
        rowgroup_sum_kernel<<<gy, NTHREADS, 0, r.s.st>>>(r.s.partial + off, gx, r.rs_d + (size_t)slot * r.rs_stride);
        r.s.launches++;
        cudaMemcpyAsync(r.rs_h + (size_t)slot * r.rs_stride, r.rs_d + (size_t)slot * r.rs_stride, sizeof(double) * gy,
                        cudaMemcpyDeviceToHost, r.s.st);
    };
    auto host_sum = [&](BatchRoom &r, int gy, int slot) {
        double s = 0.0;
        const double *p = r.rs_h + (size_t)slot * r.rs_stride;
        for (int i = 0; i < gy; i++) s += p[i];  // fixed global order
        return s;
    };
    auto lower_index = [&](double t_new) {  // ivp.py:715-728, direction < 0: first ascending index with t_eval >= t_new
        int lo = 0, hi = nt;
        while (lo < hi) {
            int mid = (lo + hi) / 2;
            if (t_eval[nt - 1 - mid] < t_new) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    // Batched launches: the step attempts that become ready during one polling sweep are grouped by their number of
    // dense-output samples NE and launched together, up to fused::BATCH_MAX rooms per launch (blockIdx.z = room), instead
    // of one launch per room: 512^2 rooms need ~10 us of device work per attempt, less than a launch round trip.  Only
    // with phi output (velocity-only storage converts scratch slices on the room's own stream) and with the in-kernel
    // final reduction; a room's arithmetic is unchanged (same kernel body, same chunking, same reduction order).
    const bool batched = d_phi != nullptr && final_in_kernel && !getenv("OC_BATCH_PER_ROOM");
    std::vector<BatchRoom *> buckets[fused::NE_MAX + 1];
    std::vector<fused::BatchArgs> host_ba(1);
    int bucket_launches = 0;
    auto flush = [&]() -> int {
        for (int ne = 0; ne <= fused::NE_MAX; ne++) {
            auto &bk = buckets[ne];
            for (size_t off = 0; off < bk.size(); off += fused::BATCH_MAX) {
                const int nb = (int)std::min<size_t>(fused::BATCH_MAX, bk.size() - off);
                for (int q = 0; q < nb; q++) {
                    BatchRoom &r = *bk[off + q];
                    fused::set_tensor_maps(r.fa, Ny, &r.s.maps);
                    host_ba[0].a[q] = r.fa;
                }
                cudaStream_t st = ctx->batch_streams[(bucket_launches++) % n_streams];
                OC_CUDA(fused::launch_batch(ne, host_ba[0], nb, dim3(fgx, fgy), st));
                rooms[0].s.launches++;
            }
            bk.clear();
        }
        return OC_OK;
    };
    // enqueue one step attempt of room r (fused kernel + reduction); returns false when the room has finished
    auto attempt = [&](BatchRoom &r) -> int {
        if (r.h_abs < r.min_step) { r.status = -1; r.phase = BatchRoom::DONE; return OC_OK; }
        double h = r.h_abs * direction;
        double t_new = r.t + h;
        if (direction * (t_new - t_bound) > 0) t_new = t_bound;
        h = t_new - r.t;
        r.h = h; r.t_new = t_new; r.h_abs = std::fabs(h);
        r.ia_lo = std::min(lower_index(t_new), r.t_eval_i);
        fused::Args fa{};
        Solver &s = r.s;
        fa.y = s.y; fa.k1 = s.K[0]; fa.coef = s.coef; fa.ynew = s.ynew; fa.k7 = s.K[6]; fa.partial = s.partial;
        fa.ha21 = h * RK_A[1][0];
        for (int j = 0; j < 2; j++) fa.ha3[j] = h * RK_A[2][j];
        for (int j = 0; j < 3; j++) fa.ha4[j] = h * RK_A[3][j];
        for (int j = 0; j < 4; j++) fa.ha5[j] = h * RK_A[4][j];
        for (int j = 0; j < 5; j++) fa.ha6[j] = h * RK_A[5][j];
        for (int j = 0; j < 6; j++) fa.hb[j] = h * RK_B[j];
        for (int j = 0; j < 7; j++) fa.he[j] = h * RK_E[j];
        fa.A = s.diff_over_dxdy; fa.rtol = rtol; fa.atol = atol;
        fa.Ny = Ny; fa.Nx = Nx; fa.RC = frc;
        fa.row_base = 0; fa.own0 = 0; fa.own1 = Ny; fa.phi_row_base = 0;
        // the samples this attempt emits if accepted, NE_MAX per launch; with more than NE_MAX samples inside one
        // step (only when h >> dt) the step is simply launched again for the remaining samples: the re-launch
        // recomputes identical y_new / f_new / error sums
        int ia = r.t_eval_i - 1;
        do {
            int ne = 0;
            for (; ia >= r.ia_lo && ne < fused::NE_MAX; ia--, ne++) {
                const int kd = nt - 1 - ia;
                const double x = (t_eval[kd] - r.t) / h;  // RkDenseOutput: (t - t_old)/h
                const double pw[4] = {x, x * x, x * x * x, x * x * x * x};
                for (int j = 0; j < 7; j++) {
                    double acc = 0.0;
                    for (int q = 0; q < 4; q++) acc += RK_P[j][q] * pw[q];
                    fa.w[ne][j] = h * acc;
                }
                // without a phi output the samples go to the room's scratch slices (velocities are derived on accept)
                fa.phi[ne] = d_phi ? r.phi + (size_t)kd * n : r.phi + (size_t)((r.t_eval_i - 1 - ia) % fused::NE_MAX) * n;
            }
            if (final_in_kernel) { final_reduce_args(ctx, (int)(&r - rooms.data()), fa); r.wait_seq = fa.seq; }
            if (batched && ia < r.ia_lo) {  // all samples fit one launch (the rule): wait in the NE bucket
                r.fa = fa;
                buckets[ne].push_back(&r);
                break;
            }
            int rc = s.launch_fused(ne, fa);
            if (rc) return rc;
            if (!d_phi && ia >= r.ia_lo) {
                oc::set_error("batched solve without a phi output supports at most %d samples per step", fused::NE_MAX);
                return OC_ERR_ARG;
            }
        } while (ia >= r.ia_lo);
        r.st.nfev += 6;
        if (!final_in_kernel) {
            reduce_async(r, 0, fgx, fgy, 0);
            cudaEventRecord(r.ev, s.st);
        }
        r.phase = BatchRoom::STEP;
        return OC_OK;
    };
    auto start_step = [&](BatchRoom &r) -> int {  // base.py:179-210 step(): one call of _step_impl
        if (r.t == t_bound) { r.status = 0; r.phase = BatchRoom::DONE; return OC_OK; }
        r.min_step = 10 * std::fabs(std::nextafter(r.t, direction * INFINITY) - r.t);
        if (r.h_abs < r.min_step) r.h_abs = r.min_step;
        r.rejected = false;
        return attempt(r);
    };
    auto advance = [&](BatchRoom &r) -> int {
        Solver &s = r.s;
        switch (r.phase) {
        case BatchRoom::START: {
            prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s.st>>>(r.V, r.m, prm->g, 1.0 / (prm->mu * s2), n, s.coef, s.y);
            s.launches++;
            Comb c{};
            c.y = s.y;
            s.stage(0, c, s.K[0]);  // rk.py:94
            r.st.nfev++;
            if (std::fabs(t_bound - T) == 0.0) { r.h_abs = 0.0; r.st.h0 = 0.0; return start_step(r); }
            init_norm_kernel<<<dim3(nbx, nby), NTHREADS, 0, s.st>>>(s.y, s.K[0], nullptr, rtol, atol, Ny, Nx, s.partial, 0);
            s.launches++;
            reduce_async(r, 0, nbx, nby, 0);
            reduce_async(r, (size_t)nbx * nby, nbx, nby, 1);
            cudaEventRecord(r.ev, s.st);
            r.phase = BatchRoom::INIT0;
            return OC_OK;
        }
        case BatchRoom::INIT0: {  // common.py:109-127
            const double d0 = std::sqrt(host_sum(r, nby, 0)) / sqrt_n, d1 = std::sqrt(host_sum(r, nby, 1)) / sqrt_n;
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            h0 = std::min(h0, std::fabs(t_bound - T));
            r.h0 = h0; r.d1 = d1;
            Comb c{};
            c.y = s.y; c.k[0] = s.K[0]; c.a[0] = 1.0; c.h = h0 * direction;
            s.stage(1, c, s.K[1]);
            r.st.nfev++;
            init_norm_kernel<<<dim3(nbx, nby), NTHREADS, 0, s.st>>>(s.y, s.K[0], s.K[1], rtol, atol, Ny, Nx, s.partial, 1);
            s.launches++;
            reduce_async(r, 0, nbx, nby, 2);
            cudaEventRecord(r.ev, s.st);
            r.phase = BatchRoom::INIT1;
            return OC_OK;
        }
        case BatchRoom::INIT1: {  // common.py:127-134
            const double d2 = std::sqrt(host_sum(r, nby, 2)) / sqrt_n / r.h0;
            double h1;
            if (r.d1 <= 1e-15 && d2 <= 1e-15) h1 = std::max(1e-6, r.h0 * 1e-3);
            else h1 = std::pow(0.01 / std::max(r.d1, d2), 1.0 / 5.0);
            r.h_abs = std::min(std::min(100 * r.h0, h1), std::fabs(t_bound - T));
            r.st.h0 = r.h_abs;
            return start_step(r);
        }
        case BatchRoom::STEP: {  // rk.py:146-165
            const double error_norm = std::sqrt(final_in_kernel ? r.step_sum : host_sum(r, fgy, 0)) / sqrt_n;
            if (!(error_norm < 1)) {
                r.h_abs *= std::max(0.2, 0.9 * std::pow(error_norm, error_exponent));
                r.rejected = true;
                r.st.n_rejected++;
                return attempt(r);
            }
            double factor = (error_norm == 0) ? 10.0 : std::min(10.0, 0.9 * std::pow(error_norm, error_exponent));
            if (r.rejected) factor = std::min(1.0, factor);
            r.h_abs *= factor;
            r.st.n_accepted++;
            const int t_eval_i_new = lower_index(r.t_new);
            for (int ia = r.t_eval_i - 1; ia >= t_eval_i_new; ia--) {
                const int kd = nt - 1 - ia;
                r.n_out++;
                if (want_v && kd >= 1) {  // slice s = vels(sol.y[:, nt-1-s]) (optimals.py:200-204)
                    const double *ph = d_phi ? r.phi + (size_t)kd * n : r.phi + (size_t)((r.t_eval_i - 1 - ia) % fused::NE_MAX) * n;
                    const int sl = nt - 1 - kd;
                    dim3 grid((Nx - 2 + 63) / 64, (Ny - 2 + 3) / 4);
                    vels_kernel<<<grid, NTHREADS, 0, s.st>>>(ph, Ny, Nx, prm->mu, prm->lim, 1.0 / (2 * ctx->dx), 1.0 / (2 * ctx->dy),
                                                             r.vx + (size_t)sl * n_int, r.vy + (size_t)sl * n_int);
                    s.launches++;
                }
            }
            r.t_eval_i = t_eval_i_new;
            const bool finished = direction * (r.t_new - t_bound) >= 0;  // base.py:205-208
            r.t = r.t_new;
            std::swap(s.y, s.ynew);
            std::swap(s.K[0], s.K[6]);
            if (finished) { r.status = 0; r.phase = BatchRoom::DONE; return OC_OK; }
            return start_step(r);
        }
        default: return OC_OK;
        }
    };
    int rc = OC_OK, live = n_rooms;
    for (int b = 0; b < n_rooms && !rc; b++) {
        rc = advance(rooms[b]);
        if (rooms[b].phase == BatchRoom::DONE) live--;
    }
    if (!rc) rc = flush();
    unsigned long long idle = 0;
    while (live > 0 && !rc) {
        bool progressed = false;
        for (int b = 0; b < n_rooms && !rc; b++) {
            BatchRoom &r = rooms[b];
            if (r.phase == BatchRoom::DONE) continue;
            if (final_in_kernel && r.phase == BatchRoom::STEP) {
                // the attempt's last CTA publishes the error sum in mapped host memory: no event, no copy
                if (!final_ready(ctx, b, r.wait_seq, &r.step_sum)) continue;
            } else {
                cudaError_t e = cudaEventQuery(r.ev);
                if (e == cudaErrorNotReady) continue;
                OC_CUDA(e);
            }
            progressed = true;
            rc = advance(r);
            if (r.phase == BatchRoom::DONE) live--;
        }
        if (!rc) rc = flush();  // one batched launch per NE class for everything that became ready in this sweep
        if (progressed) { idle = 0; continue; }
        if ((++idle & 0x3ffff) == 0) {  // nothing became ready for a long time: make sure no launch has failed
            for (int q = 0; q < n_streams && !rc; q++) {
                cudaError_t e = cudaStreamQuery(ctx->batch_streams[q]);
                if (e != cudaSuccess && e != cudaErrorNotReady) {
                    oc::set_error("batched HJB solve: %s", cudaGetErrorString(e));
                    cudaGetLastError();
                    rc = OC_ERR_CUDA;
                }
            }
        }
    }
    // the caller's stream continues after every room stream
    for (int q = 0; q < n_streams; q++) {
        cudaEventRecord(ctx->batch_events[q], ctx->batch_streams[q]);
        cudaStreamWaitEvent(main_st, ctx->batch_events[q], 0);
    }
    OC_CUDA(cudaEventRecord(ctx->ev1, main_st));
    OC_CUDA(cudaStreamSynchronize(main_st));
    OC_CUDA(cudaGetLastError());
    if (rc) return rc;
    float ms = 0;
    OC_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    int bad = -1, total_launches = 0;
    for (int b = 0; b < n_rooms; b++) {
        BatchRoom &r = rooms[b];
        r.st.status = r.status; r.st.n_out = r.n_out; r.st.launches = r.s.launches; r.st.gpu_ms = ms;
        total_launches += r.s.launches;
        stats[b] = r.st;
        if (r.status == -1 && bad < 0) bad = b;
    }
    oc::count_launch(total_launches);
    if (bad >= 0) {
        oc::set_error("RK45: required step size is less than spacing between numbers (room %d, t=%g)", bad, rooms[bad].t);
        return OC_ERR_STEP_TOO_SMALL;
    }
    return OC_OK;
}
