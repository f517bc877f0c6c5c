// Per-environment algorithms of the engine, written once and compiled twice:
//   * by nvcc for sm_100a (the product: one CTA per environment, working set in
//     shared memory, warp-shuffle reductions) and
//   * by g++ with -DOPFG_HOSTSIM into tests/hostsim/libopfg_hostsim.so, where one
//     host "thread" (tid 0 of 1) walks the same tables.  The host build exists so
//     that the symbolic schedule and the table marshalling can be unit-tested on
//     the GPU-less builder box; it is NOT reachable from the product package.
#pragma once
#include <cmath>
#include <cstdint>

#include "../../include/opfg_b200.h"

#if defined(__CUDACC__) && !defined(OPFG_HOSTSIM)
#define OPFG_DEVICE_BUILD 1
#define OPFG_HD __device__ __forceinline__
#define OPFG_HHD __host__ __device__ __forceinline__
#else
#define OPFG_HD inline
#define OPFG_HHD inline
#endif

namespace opfg {

struct GridDev {
    // sizes
    int nb, n, n_levels, n_blocks, n_fill, nnz_y, nbr, ng, n_ref;
    int threads;
    double base_mva, tol;
    int max_iter, init_dc;
    // numbering
    const int* bus_of_int;
    const int* int_of_bus;
    const unsigned char* type_int;     // bus type by internal index
    const double* vm0_int;             // start |V| (init_vm_pu, or the set-point at generator buses)
    const double* va0_int;             // start angle (rad); ref buses keep it
    // schedule
    const int* level_ptr;
    const int* fill_ids;
    const int *dp_ptr, *dp_l, *dp_w, *dp_m;
    const int *off_ptr, *off_tgt, *off_piv, *op_ptr, *op_l, *op_w;
    const int *up_ptr, *up_w, *up_j;
    // Ybus
    const int *y_ptr, *y_col, *y_blk, *y_diag;
    const double* y_val;               // [nnz_y*2] re,im  (written by the Ybus assembly kernel)
    const int *yc_ptr, *yc_branch, *yc_role;
    const double* br_param;            // [nbr*6] r x b g tap shift_rad  (ppc branch table)
    double* br_y;                      // [nbr*8] Yff Yft Ytf Ytt (re,im)
    const double* bus_ysh;             // [nb*2] (GS + jBS)/base by ppc bus
    const int* br_f;                   // [nbr] ppc from bus
    const int* br_t;
    // DC start
    const double* dc_val;              // [n_blocks] scalar factor on the same schedule
    const double* dc_rhs0;             // [n]
    // ---- assembly (kernel 1) ----
    int n_state, n_const, n_act, n_inj;
    const double* consts;
    const int* act_slot;
    const int *act_lo, *act_hi, *act_div, *act_kind, *act_clamp_lo, *act_clamp_hi;
    const int* inj_ptr;                // [nb+1] CSR by ppc bus
    const int *inj_p, *inj_q, *inj_coef;
    // ---- scoring (kernel 5) ----
    int n_pp_bus, res_vm_slot, res_va_slot;
    const int* pp_lookup;
    const int *br_loading_slot, *br_flow_slot;
    const double *rate_f, *rate_t;
    const int *gen_bus, *gen_p_slot, *gen_q_slot;
    const double* gen_q_share;
    int n_con;
    const int* con_ptr;
    const int *con_value, *con_min, *con_max;
    const double *con_value_scale, *con_bound_mul, *con_autoscale, *con_pfactor, *con_ppower, *con_pcount;
    const int* con_worst;
    int n_poly, n_pwl, n_pwl_seg;
    const int *poly_p, *poly_q, *poly_coef, *pwl_v, *pwl_seg;
    const double *poly_p_mul, *poly_q_mul, *pwl_v_mul;
    int reward_kind;
    double penalty_weight, clip_lo, clip_hi, obj_factor, obj_bias, pen_factor, pen_bias;
    double valid_reward, invalid_penalty, invalid_obj_share;
    int n_obs;
    const int* obs_ref;
};

// ------------------------------------------------------------------ block context
#ifdef OPFG_DEVICE_BUILD
template <int T>
struct Ctx {
    int tid;
    double* red;   // shared scratch, >= T/32 doubles
    __device__ __forceinline__ int nthreads() const { return T; }
    __device__ __forceinline__ void sync() const {
        if (T == 32) __syncwarp(); else __syncthreads();
    }
    __device__ __forceinline__ double warp_max(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ __forceinline__ double warp_sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    // NaN-propagating max over the block (fmax would drop NaNs: carry a flag)
    __device__ __forceinline__ double block_max(double v) const {
        double bad = (v != v) ? 1.0 : 0.0;
        v = warp_max(bad > 0 ? 0.0 : v);
        bad = warp_max(bad);
        if (T > 32) {
            __syncthreads();
            if ((tid & 31) == 0) { red[tid >> 5] = v; red[(T >> 5) + (tid >> 5)] = bad; }
            __syncthreads();
            v = red[0]; bad = red[T >> 5];
#pragma unroll
            for (int w = 1; w < (T >> 5); ++w) { v = fmax(v, red[w]); bad = fmax(bad, red[(T >> 5) + w]); }
            __syncthreads();
        }
        return bad > 0 ? NAN : v;
    }
    __device__ __forceinline__ double block_sum(double v) const {
        v = warp_sum(v);
        if (T > 32) {
            __syncthreads();
            if ((tid & 31) == 0) red[tid >> 5] = v;
            __syncthreads();
            v = 0;
#pragma unroll
            for (int w = 0; w < (T >> 5); ++w) v += red[w];
            __syncthreads();
        }
        return v;
    }
};
#else
template <int T>
struct Ctx {
    int tid = 0;
    double* red = nullptr;
    int nthreads() const { return 1; }
    void sync() const {}
    double block_max(double v) const { return v; }
    double block_sum(double v) const { return v; }
};
#endif

OPFG_HD double ref_val(const GridDev& g, const double* S, int r) { return r >= 0 ? S[r] : g.consts[-r - 1]; }

// ------------------------------------------------- kernel 1a: branch -> Ybus values
// Admittances of one branch from its ppc row (pypower makeYbus.py [ext-mem]).
OPFG_HD void branch_admittance(const double* p, double* y) {
    const double r = p[0], x = p[1], b = p[2], gsh = p[3];
    double tap = p[4];
    const double sh = p[5];
    if (tap == 0.0) tap = 1.0;
    const double z2 = r * r + x * x;
    const double ysr = r / z2, ysi = -x / z2;
    const double ttr = ysr + 0.5 * gsh, tti = ysi + 0.5 * b;
    const double c = cos(sh), s = sin(sh);
    const double t2 = tap * tap;
    y[0] = ttr / t2;  y[1] = tti / t2;                        // Yff = Ytt / |tap|^2
    // Yft = -Ys / conj(tap) = -Ys * tap / |tap|^2 ; tap = tap*(c + js)
    y[2] = -(ysr * c - ysi * s) / tap;  y[3] = -(ysr * s + ysi * c) / tap;
    // Ytf = -Ys / tap = -Ys * conj(tap) / |tap|^2
    y[4] = -(ysr * c + ysi * s) / tap;  y[5] = -(-ysr * s + ysi * c) / tap;
    y[6] = ttr;  y[7] = tti;
}

// one Ybus CSR entry = ordered sum of its branch / shunt contributions
OPFG_HD void ybus_entry(const GridDev& g, const double* br_y, int e, double* out) {
    double re = 0, im = 0;
    for (int c = g.yc_ptr[e]; c < g.yc_ptr[e + 1]; ++c) {
        const int role = g.yc_role[c], idx = g.yc_branch[c];
        if (role == 4) { re += g.bus_ysh[2 * idx]; im += g.bus_ysh[2 * idx + 1]; }
        else { re += br_y[8 * idx + 2 * role]; im += br_y[8 * idx + 2 * role + 1]; }
    }
    out[0] = re; out[1] = im;
}

// --------------------------------------- kernel 1b: actions -> set-points -> Sbus
template <class C>
OPFG_HD void env_assemble(const GridDev& g, const C& cx, const double* act, double* S, double* sbus) {
    const int T = cx.nthreads();
    for (int j = cx.tid; act != nullptr && j < g.n_act; j += T) {
        double a = act[j];
        a = a < 0.0 ? 0.0 : (a > 1.0 ? 1.0 : a);               // opf_env.py:429
        const double lo = ref_val(g, S, g.act_lo[j]), hi = ref_val(g, S, g.act_hi[j]);
        double sp = a * (hi - lo) + lo;                           // :461
        if (g.act_clamp_lo) {                                     // :464-470
            const double cl = ref_val(g, S, g.act_clamp_lo[j]), ch = ref_val(g, S, g.act_clamp_hi[j]);
            if (sp > ch) sp = ch;
            if (sp < cl) sp = cl;
        }
        sp /= ref_val(g, S, g.act_div[j]);                        // :472-474
        const int kind = g.act_kind[j];
        if (kind == 1) sp = (rint(sp) != 0.0) ? 1.0 : 0.0;        // :476-478
        else if (kind == 2) sp = rint(sp);                        // :479-481
        S[g.act_slot[j]] = sp;
    }
    cx.sync();
#ifdef OPFG_DEVICE_BUILD
    __threadfence_block();
#endif
    const double inv_base = 1.0 / g.base_mva;
    for (int bus = cx.tid; bus < g.nb; bus += T) {
        double p = 0, q = 0;
        for (int e = g.inj_ptr[bus]; e < g.inj_ptr[bus + 1]; ++e) {
            const double c = ref_val(g, S, g.inj_coef[e]);
            p += c * ref_val(g, S, g.inj_p[e]);
            q += c * ref_val(g, S, g.inj_q[e]);
        }
        sbus[2 * bus] = p * inv_base;
        sbus[2 * bus + 1] = q * inv_base;
    }
}

// ------------------------------------------------------ kernels 2-4: Newton-Raphson
struct PfSmem {
    double *lu, *rhs, *vr, *vi, *vm, *va, *psp, *qsp, *red;
};

OPFG_HHD size_t pf_smem_doubles(int n_blocks, int n, int nb, int threads) {
    return (size_t)4 * n_blocks + 2 * (size_t)n + 6 * (size_t)nb + 2 * (size_t)(threads / 32 + 1);
}

OPFG_HD PfSmem pf_carve(double* base, int n_blocks, int n, int nb) {
    PfSmem s;
    s.lu = base;
    s.rhs = s.lu + 4 * (size_t)n_blocks;
    s.vr = s.rhs + 2 * (size_t)n;
    s.vi = s.vr + nb;
    s.vm = s.vi + nb;
    s.va = s.vm + nb;
    s.psp = s.va + nb;
    s.qsp = s.psp + nb;
    s.red = s.qsp + nb;
    return s;
}

// Fused power mismatch (kernel 2) + Jacobian assembly into the fixed block
// pattern (kernel 3) for block row i.  Returns the row's contribution to ||F||inf.
// Formulas: pypower dSbus_dV.py in polar form [ext-mem], SURVEY.md App. B.4.
OPFG_HD double row_mismatch_jacobian(const GridDev& g, const PfSmem& s, const double* yv, int i) {
    const double vir = s.vr[i], vii = s.vi[i], vmi = s.vm[i];
    const bool pq = g.type_int[i] == OPFG_PQ;
    double ir = 0, ii = 0, diag_ar = 0, diag_ai = 0;
    const int e0 = g.y_ptr[i], e1 = g.y_ptr[i + 1];
    for (int e = e0; e < e1; ++e) {
        const int j = g.y_col[e];
        const double gr = yv[2 * e], bi = yv[2 * e + 1];
        const double vjr = s.vr[j], vji = s.vi[j];
        const double tr = gr * vjr - bi * vji, ti = gr * vji + bi * vjr;   // Y_ij V_j
        ir += tr; ii += ti;
        const double ar = vir * tr + vii * ti, ai = vii * tr - vir * ti;   // V_i conj(Y_ij V_j)
        if (e == e0) { diag_ar = ar; diag_ai = ai; continue; }             // diagonal entry is first
        const int blk = g.y_blk[e];
        if (blk < 0) continue;                                             // column is a ref bus
        const double inv_vmj = 1.0 / s.vm[j];
        double* b = s.lu + 4 * (size_t)blk;
        b[0] = ai;                 // dP_i/dtheta_j
        b[1] = ar * inv_vmj;       // dP_i/dVm_j
        b[2] = pq ? -ar : 0.0;     // dQ_i/dtheta_j
        b[3] = pq ? ai * inv_vmj : 0.0;
    }
    const double P = vir * ir + vii * ii, Q = vii * ir - vir * ii;          // S_i = V_i conj(I_i)
    const double inv_vmi = 1.0 / vmi;
    double* d = s.lu + 4 * (size_t)i;
    d[0] = -Q + diag_ai;
    d[1] = (diag_ar + P) * inv_vmi;
    d[2] = pq ? P - diag_ar : 0.0;
    d[3] = pq ? (diag_ai + Q) * inv_vmi : 1.0;
    const double dp = P - s.psp[i], dq = pq ? Q - s.qsp[i] : 0.0;
    s.rhs[2 * i] = -dp;
    s.rhs[2 * i + 1] = -dq;
    const double a = fabs(dp), c = fabs(dq);
    if (dp != dp || dq != dq) return NAN;
    return a > c ? a : c;
}

OPFG_HD void lu_diag_item(const GridDev& g, const PfSmem& s, int k) {
    double* D = s.lu + 4 * (size_t)k;
    double a = D[0], b = D[1], c = D[2], d = D[3];
    double y0 = s.rhs[2 * k], y1 = s.rhs[2 * k + 1];
    for (int p = g.dp_ptr[k]; p < g.dp_ptr[k + 1]; ++p) {
        const double* L = s.lu + 4 * (size_t)g.dp_l[p];
        const double* W = s.lu + 4 * (size_t)g.dp_w[p];
        const int m = g.dp_m[p];
        const double l0 = L[0], l1 = L[1], l2 = L[2], l3 = L[3];
        const double w0 = W[0], w1 = W[1], w2 = W[2], w3 = W[3];
        const double t0 = s.rhs[2 * m], t1 = s.rhs[2 * m + 1];
        a -= l0 * w0 + l1 * w2;  b -= l0 * w1 + l1 * w3;
        c -= l2 * w0 + l3 * w2;  d -= l2 * w1 + l3 * w3;
        y0 -= l0 * t0 + l1 * t1; y1 -= l2 * t0 + l3 * t1;
    }
    const double r = 1.0 / (a * d - b * c);
    const double ia = d * r, ib = -b * r, ic = -c * r, id = a * r;
    D[0] = ia; D[1] = ib; D[2] = ic; D[3] = id;
    s.rhs[2 * k] = ia * y0 + ib * y1;
    s.rhs[2 * k + 1] = ic * y0 + id * y1;
}

OPFG_HD void lu_off_item(const GridDev& g, const PfSmem& s, int item) {
    double* X = s.lu + 4 * (size_t)g.off_tgt[item];
    double a = X[0], b = X[1], c = X[2], d = X[3];
    for (int p = g.op_ptr[item]; p < g.op_ptr[item + 1]; ++p) {
        const double* L = s.lu + 4 * (size_t)g.op_l[p];
        const double* W = s.lu + 4 * (size_t)g.op_w[p];
        const double l0 = L[0], l1 = L[1], l2 = L[2], l3 = L[3];
        const double w0 = W[0], w1 = W[1], w2 = W[2], w3 = W[3];
        a -= l0 * w0 + l1 * w2;  b -= l0 * w1 + l1 * w3;
        c -= l2 * w0 + l3 * w2;  d -= l2 * w1 + l3 * w3;
    }
    const int piv = g.off_piv[item];
    if (piv >= 0) {   // W = D^-1 * U
        const double* I = s.lu + 4 * (size_t)piv;
        const double i0 = I[0], i1 = I[1], i2 = I[2], i3 = I[3];
        const double na = i0 * a + i1 * c, nb_ = i0 * b + i1 * d;
        const double nc = i2 * a + i3 * c, nd = i2 * b + i3 * d;
        a = na; b = nb_; c = nc; d = nd;
    }
    X[0] = a; X[1] = b; X[2] = c; X[3] = d;
}

OPFG_HD void bwd_item(const GridDev& g, const PfSmem& s, int k) {
    double x0 = s.rhs[2 * k], x1 = s.rhs[2 * k + 1];
    for (int p = g.up_ptr[k]; p < g.up_ptr[k + 1]; ++p) {
        const double* W = s.lu + 4 * (size_t)g.up_w[p];
        const int j = g.up_j[p];
        const double xj0 = s.rhs[2 * j], xj1 = s.rhs[2 * j + 1];
        x0 -= W[0] * xj0 + W[1] * xj1;
        x1 -= W[2] * xj0 + W[3] * xj1;
    }
    s.rhs[2 * k] = x0;
    s.rhs[2 * k + 1] = x1;
}

// One environment: DC start, Newton-Raphson to tolerance, write |V|, angle, flag.
// Mirrors pandapower newtonpf control flow: convergence is
// tested before the first iteration; at most max_iter linear solves.
template <class C>
OPFG_HD void env_pf_solve(const GridDev& g, const C& cx, double* smem, const double* sbus,
                          const double* yval_env, double* vm_out, double* va_out,
                          uint8_t* conv_out, int32_t* iter_out) {
    const int T = cx.nthreads();
    const int n = g.n, nb = g.nb;
    PfSmem s = pf_carve(smem, g.n_blocks, n, nb);
    const double* yv = yval_env ? yval_env : g.y_val;

    for (int i = cx.tid; i < nb; i += T) {
        const int bus = g.bus_of_int[i];
        s.psp[i] = sbus[2 * bus];
        s.qsp[i] = sbus[2 * bus + 1];
        s.vm[i] = g.vm0_int[i];
        s.va[i] = g.va0_int[i];
    }
    cx.sync();

    if (g.init_dc) {   // pandapower init='dc': B' theta = P on the shared, pre-factorised B'
        for (int k = cx.tid; k < n; k += T) s.rhs[k] = s.psp[k] + g.dc_rhs0[k];
        cx.sync();
        for (int l = 0; l < g.n_levels; ++l) {
            for (int k = g.level_ptr[l] + cx.tid; k < g.level_ptr[l + 1]; k += T) {
                double y = s.rhs[k];
                for (int p = g.dp_ptr[k]; p < g.dp_ptr[k + 1]; ++p) y -= g.dc_val[g.dp_l[p]] * s.rhs[g.dp_m[p]];
                s.rhs[k] = y * g.dc_val[k];
            }
            cx.sync();
        }
        for (int l = g.n_levels - 1; l >= 0; --l) {
            for (int k = g.level_ptr[l] + cx.tid; k < g.level_ptr[l + 1]; k += T) {
                double x = s.rhs[k];
                for (int p = g.up_ptr[k]; p < g.up_ptr[k + 1]; ++p) x -= g.dc_val[g.up_w[p]] * s.rhs[g.up_j[p]];
                s.rhs[k] = x;
            }
            cx.sync();
        }
        for (int k = cx.tid; k < n; k += T) s.va[k] = s.rhs[k];
        cx.sync();
    }
    for (int i = cx.tid; i < nb; i += T) {
        double sn, cs;
        sincos(s.va[i], &sn, &cs);
        s.vr[i] = s.vm[i] * cs;
        s.vi[i] = s.vm[i] * sn;
    }
    cx.sync();

    int it = 0;
    int converged = 0;
    while (true) {
        for (int f = cx.tid; f < g.n_fill; f += T) {
            double* b = s.lu + 4 * (size_t)g.fill_ids[f];
            b[0] = 0; b[1] = 0; b[2] = 0; b[3] = 0;
        }
        double nrm = 0;
        bool bad = false;
        for (int i = cx.tid; i < n; i += T) {
            const double r = row_mismatch_jacobian(g, s, yv, i);
            if (r != r) bad = true; else if (r > nrm) nrm = r;
        }
        nrm = cx.block_max(bad ? NAN : nrm);
        cx.sync();
        if (nrm < g.tol) { converged = 1; break; }
        if (it >= g.max_iter || nrm != nrm) break;
        ++it;
        int item = 0;
        for (int l = 0; l < g.n_levels; ++l) {
            for (int k = g.level_ptr[l] + cx.tid; k < g.level_ptr[l + 1]; k += T) lu_diag_item(g, s, k);
            cx.sync();
            for (item = g.off_ptr[l] + cx.tid; item < g.off_ptr[l + 1]; item += T) lu_off_item(g, s, item);
            cx.sync();
        }
        for (int l = g.n_levels - 1; l >= 0; --l) {
            for (int k = g.level_ptr[l] + cx.tid; k < g.level_ptr[l + 1]; k += T) bwd_item(g, s, k);
            cx.sync();
        }
        for (int k = cx.tid; k < n; k += T) {
            double va = s.va[k] + s.rhs[2 * k];
            double vm = s.vm[k] + ((g.type_int[k] == OPFG_PQ) ? s.rhs[2 * k + 1] : 0.0);
            // V = Vm*exp(j*Va); Vm = |V|; Va = angle(V)  (newtonpf.py)
            if (vm < 0) { vm = -vm; va += M_PI; }
            if (va > M_PI || va <= -M_PI) va -= 2.0 * M_PI * floor((va + M_PI) / (2.0 * M_PI));
            double sn, cs;
            sincos(va, &sn, &cs);
            s.va[k] = va; s.vm[k] = vm;
            s.vr[k] = vm * cs; s.vi[k] = vm * sn;
        }
        cx.sync();
    }
    for (int i = cx.tid; i < nb; i += T) {
        const int bus = g.bus_of_int[i];
        vm_out[bus] = s.vm[i];
        va_out[bus] = s.va[i];
    }
    if (cx.tid == 0) { *conv_out = (uint8_t)converged; *iter_out = it; }
}

// ------------------------------------------------------------- kernel 5: scoring
struct ScoreSmem {
    double *vr, *vi, *vm, *sf, *st, *red;
};

OPFG_HHD size_t score_smem_doubles(int nb, int nbr, int threads) {
    return 3 * (size_t)nb + 4 * (size_t)nbr + 2 * (size_t)(threads / 32 + 1);
}

OPFG_HD double pwl_cost(const GridDev& g, const double* S, int row, double v) {
    // opfgym/objective.py:57-77; `beyond` has no sign test
    const double sg = (v > 0) - (v < 0);
    const double mag = fabs(v);
    double total = 0;
    for (int k = 0; k < g.n_pwl_seg; ++k) {
        const int* r = g.pwl_seg + 3 * ((size_t)row * g.n_pwl_seg + k);
        const double lo = ref_val(g, S, r[0]), hi = ref_val(g, S, r[1]), price = ref_val(g, S, r[2]);
        const double alo = fabs(lo), ahi = fabs(hi);
        const double near_ = alo < ahi ? alo : ahi, far_ = alo < ahi ? ahi : alo;
        const double ssum = lo + hi;
        const double sref = (ssum > 0) - (ssum < 0);
        const bool beyond = mag > far_;
        const bool within = (mag > near_) && (sg == sref) && !beyond && (v == v);
        if (beyond) total += sg * (hi - lo) * price;
        if (within) total += sg * (mag - near_) * price;
    }
    return total;
}

template <class C>
OPFG_HD void env_score(const GridDev& g, const C& cx, double* smem, const OpfgBatch& B, int64_t env,
                       const double* yval_env) {
    const int T = cx.nthreads();
    const int nb = g.nb, nbr = g.nbr, nc = g.n_con;
    ScoreSmem s;
    s.vr = smem; s.vi = s.vr + nb; s.vm = s.vi + nb;
    s.sf = s.vm + nb; s.st = s.sf + 2 * (size_t)nbr; s.red = s.st + 2 * (size_t)nbr;
    double* S = B.state + env * (int64_t)g.n_state;
    const double* vm = B.vm + env * (int64_t)nb;
    const double* va = B.va + env * (int64_t)nb;
    const double* sbus = B.sbus + env * (int64_t)nb * 2;
    const bool conv = B.converged[env] != 0;
    const double base = g.base_mva;
    const double* yv = yval_env ? yval_env : g.y_val;

    if (!conv) {
        // opf_env.py:390-399: NaN observation and reward, every constraint reported violated
        for (int j = cx.tid; j < g.n_obs; j += T) {
            if (B.obs_f32) B.obs_f32[env * (int64_t)g.n_obs + j] = NAN;
            if (B.obs_f64) B.obs_f64[env * (int64_t)g.n_obs + j] = NAN;
        }
        for (int c = cx.tid; c < nc; c += T) {
            if (B.valids) B.valids[env * nc + c] = 0;
            if (B.violations) B.violations[env * nc + c] = 1.0;
            if (B.penalties) B.penalties[env * nc + c] = 1.0;
        }
        if (cx.tid == 0) {
            if (B.reward) B.reward[env] = NAN;
            if (B.objective) B.objective[env] = NAN;
            if (B.penalty) B.penalty[env] = (double)nc;
            if (B.cost) B.cost[env] = NAN;
#ifdef OPFG_DEVICE_BUILD
            if (B.stats) atomicAdd(B.stats + OPFG_STAT_N, 1.0);
#else
            if (B.stats) B.stats[OPFG_STAT_N] += 1.0;
#endif
        }
        return;
    }

    for (int i = cx.tid; i < nb; i += T) {
        double sn, cs;
        sincos(va[i], &sn, &cs);
        s.vm[i] = vm[i];
        s.vr[i] = vm[i] * cs;
        s.vi[i] = vm[i] * sn;
    }
    cx.sync();
    // branch flows and loading (pfsoln + results_branch.py [ext-mem])
    for (int l = cx.tid; l < nbr; l += T) {
        const double* y = g.br_y + 8 * (size_t)l;
        const int f = g.br_f[l], t = g.br_t[l];
        const double vfr = s.vr[f], vfi = s.vi[f], vtr = s.vr[t], vti = s.vi[t];
        const double ifr = y[0] * vfr - y[1] * vfi + y[2] * vtr - y[3] * vti;
        const double ifi = y[0] * vfi + y[1] * vfr + y[2] * vti + y[3] * vtr;
        const double itr = y[4] * vfr - y[5] * vfi + y[6] * vtr - y[7] * vti;
        const double iti = y[4] * vfi + y[5] * vfr + y[6] * vti + y[7] * vtr;
        const double pf = (vfr * ifr + vfi * ifi) * base, qf = (vfi * ifr - vfr * ifi) * base;
        const double pt = (vtr * itr + vti * iti) * base, qt = (vti * itr - vtr * iti) * base;
        s.sf[2 * l] = pf; s.sf[2 * l + 1] = qf; s.st[2 * l] = pt; s.st[2 * l + 1] = qt;
        const double lf = sqrt(pf * pf + qf * qf) * g.rate_f[l] / s.vm[f];
        const double lt = sqrt(pt * pt + qt * qt) * g.rate_t[l] / s.vm[t];
        const int slot = g.br_loading_slot[l];
        if (slot >= 0) S[slot] = 100.0 * (lf > lt ? lf : lt);
        const int fs = g.br_flow_slot[l];
        if (fs >= 0) { S[fs] = pf; S[fs + 1] = qf; S[fs + 2] = pt; S[fs + 3] = qt; }
    }
    // bus results in pandapower bus order (fused buses repeat their value)
    if (g.res_vm_slot >= 0)
        for (int b = cx.tid; b < g.n_pp_bus; b += T) {
            const int i = g.pp_lookup[b];
            S[g.res_vm_slot + b] = i >= 0 ? vm[i] : NAN;
            if (g.res_va_slot >= 0) S[g.res_va_slot + b] = i >= 0 ? va[i] * (180.0 / M_PI) : NAN;
        }
    // generator results: slack P, and Q of every voltage-controlled generator
    for (int gi = cx.tid; gi < g.ng; gi += T) {
        const int ps = g.gen_p_slot[gi], qs = g.gen_q_slot[gi];
        if (ps < 0 && qs < 0) continue;
        const int bus = g.gen_bus[gi];
        const int i = g.int_of_bus[bus];
        double ir = 0, ii = 0;
        for (int e = g.y_ptr[i]; e < g.y_ptr[i + 1]; ++e) {
            const int j = g.bus_of_int[g.y_col[e]];
            ir += yv[2 * e] * s.vr[j] - yv[2 * e + 1] * s.vi[j];
            ii += yv[2 * e] * s.vi[j] + yv[2 * e + 1] * s.vr[j];
        }
        const double P = s.vr[bus] * ir + s.vi[bus] * ii, Q = s.vi[bus] * ir - s.vr[bus] * ii;
        if (ps >= 0) S[ps] = (P - sbus[2 * bus]) * base;
        if (qs >= 0) S[qs] = (Q - sbus[2 * bus + 1]) * base * g.gen_q_share[gi];
    }
    cx.sync();
#ifdef OPFG_DEVICE_BUILD
    __threadfence_block();
#endif
    // constraints (opfgym/constraints.py:70-128)
    double pen_sum = 0;
    bool all_valid = true;
    for (int c = 0; c < nc; ++c) {
        double viol = 0, cnt = 0;
        const bool worst = g.con_worst[c] != 0;
        for (int e = g.con_ptr[c] + cx.tid; e < g.con_ptr[c + 1]; e += T) {
            const double v = ref_val(g, S, g.con_value[e]) * g.con_value_scale[e];
            const double mul = g.con_bound_mul[e];
            const double hi = ref_val(g, S, g.con_max[e]) * mul, lo = ref_val(g, S, g.con_min[e]) * mul;
            if (v > hi) { const double x = fabs(v - hi); viol = worst ? (x > viol ? x : viol) : viol + x; cnt += 1; }
            if (v < lo) { const double x = fabs(v - lo); viol = worst ? (x > viol ? x : viol) : viol + x; cnt += 1; }
        }
        cnt = cx.block_sum(cnt);
        viol = worst ? cx.block_max(viol) : cx.block_sum(viol);
        viol *= g.con_autoscale[c];
        const double pen = -(pow(viol, g.con_ppower[c]) * g.con_pfactor[c] + cnt * g.con_pcount[c]);
        pen_sum += pen;
        if (cnt > 0) all_valid = false;
        if (cx.tid == 0) {
            if (B.valids) B.valids[env * nc + c] = cnt > 0 ? 0 : 1;
            if (B.violations) B.violations[env * nc + c] = viol;
            if (B.penalties) B.penalties[env * nc + c] = pen;
#ifdef OPFG_DEVICE_BUILD
            if (B.stats && cnt > 0 && c < OPFG_N_STATS - OPFG_STAT_VIOLATED0) atomicAdd(B.stats + OPFG_STAT_VIOLATED0 + c, 1.0);
#else
            if (B.stats && cnt > 0 && c < OPFG_N_STATS - OPFG_STAT_VIOLATED0) B.stats[OPFG_STAT_VIOLATED0 + c] += 1.0;
#endif
        }
    }
    // objective = -sum(costs)  (opfgym/objective.py:6-87, opf_env.py:493-500,517)
    double csum = 0;
    for (int r = cx.tid; r < g.n_poly; r += T) {
        const double p = ref_val(g, S, g.poly_p[r]) * g.poly_p_mul[r];
        const double q = ref_val(g, S, g.poly_q[r]) * g.poly_q_mul[r];
        const int* cf = g.poly_coef + 6 * (size_t)r;
        csum += ref_val(g, S, cf[0]) + ref_val(g, S, cf[1]) * p + ref_val(g, S, cf[2]) * p * p;
        csum += ref_val(g, S, cf[3]) + ref_val(g, S, cf[4]) * q + ref_val(g, S, cf[5]) * q * q;
    }
    for (int r = cx.tid; r < g.n_pwl; r += T)
        csum += pwl_cost(g, S, r, ref_val(g, S, g.pwl_v[r]) * g.pwl_v_mul[r]);
    const double objective = -cx.block_sum(csum);
    // reward (opfgym/reward.py:61-98 and subclasses)
    if (cx.tid == 0) {
        double obj = objective, pen = pen_sum;
        if (g.reward_kind == OPFG_REWARD_REPLACEMENT) obj = all_valid ? obj + g.valid_reward : 0.0;
        else if (g.reward_kind == OPFG_REWARD_PARAMETERIZED) {
            pen = all_valid ? pen + g.valid_reward : pen - g.invalid_penalty;
            if (!all_valid) obj *= g.invalid_obj_share;
        } else if (g.reward_kind == OPFG_REWARD_ONLY_OBJECTIVE) pen = 0.0;
        obj = obj * g.obj_factor + g.obj_bias;
        pen = pen * g.pen_factor + g.pen_bias;
        const double w = g.penalty_weight;
        double r = (w != w) ? obj + pen : obj * (1.0 - w) + pen * w;
        if (g.clip_lo == g.clip_lo) r = r < g.clip_lo ? g.clip_lo : (r > g.clip_hi ? g.clip_hi : r);
        double cost = all_valid ? 0.0 : fabs(pen_sum * g.pen_factor);
        if (!all_valid && g.reward_kind == OPFG_REWARD_PARAMETERIZED) cost += g.invalid_penalty;
        if (B.reward) B.reward[env] = r;
        if (B.objective) B.objective[env] = objective;
        if (B.penalty) B.penalty[env] = pen_sum;
        if (B.cost) B.cost[env] = cost;
        if (B.stats) {
            const double it = B.iterations ? (double)B.iterations[env] : 0.0;
#ifdef OPFG_DEVICE_BUILD
            atomicAdd(B.stats + OPFG_STAT_N, 1.0);
            atomicAdd(B.stats + OPFG_STAT_CONVERGED, 1.0);
            if (all_valid) atomicAdd(B.stats + OPFG_STAT_VALID, 1.0);
            atomicAdd(B.stats + OPFG_STAT_SUM_REWARD, r);
            atomicAdd(B.stats + OPFG_STAT_SUM_REWARD_SQ, r * r);
            atomicAdd(B.stats + OPFG_STAT_SUM_OBJECTIVE, objective);
            atomicAdd(B.stats + OPFG_STAT_SUM_PENALTY, pen_sum);
            atomicAdd(B.stats + OPFG_STAT_SUM_ITERS, it);
#else
            B.stats[OPFG_STAT_N] += 1.0; B.stats[OPFG_STAT_CONVERGED] += 1.0;
            if (all_valid) B.stats[OPFG_STAT_VALID] += 1.0;
            B.stats[OPFG_STAT_SUM_REWARD] += r; B.stats[OPFG_STAT_SUM_REWARD_SQ] += r * r;
            B.stats[OPFG_STAT_SUM_OBJECTIVE] += objective; B.stats[OPFG_STAT_SUM_PENALTY] += pen_sum;
            B.stats[OPFG_STAT_SUM_ITERS] += it;
#endif
        }
    }
    // observation gather (opf_env.py:532-549)
    for (int j = cx.tid; j < g.n_obs; j += T) {
        const double v = ref_val(g, S, g.obs_ref[j]);
        if (B.obs_f32) B.obs_f32[env * (int64_t)g.n_obs + j] = (float)v;
        if (B.obs_f64) B.obs_f64[env * (int64_t)g.n_obs + j] = v;
    }
}

// ------------------------------------------------------------------------ Philox
OPFG_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                           uint32_t* out) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two doubles in [0,1) with 53 random bits each (numpy's uint64 -> double rule)
OPFG_HD void philox_two_doubles(uint64_t seed, uint64_t env, uint64_t stream, uint32_t pair, double* a, double* b) {
    uint32_t o[4];
    philox4x32_10(pair, (uint32_t)env, (uint32_t)(env >> 32), (uint32_t)stream ^ (uint32_t)(stream >> 32) * 0x9E3779B9u,
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const uint64_t x = ((uint64_t)o[1] << 32) | o[0], y = ((uint64_t)o[3] << 32) | o[2];
    *a = (double)(x >> 11) * (1.0 / 9007199254740992.0);
    *b = (double)(y >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace opfg
