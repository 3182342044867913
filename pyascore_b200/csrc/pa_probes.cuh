// Stage probes behind the secondary public classes of the reference (PyModifiedPeptide /
// PyFragmentGraph, PyLogMath, PyBinomialDist): small kernels that evaluate ONE object's values
// with the very device functions of the hot path, so the Python wrappers hold no arithmetic.
#pragma once
#include "pa_kernels.cuh"

// cpp/ModifiedPeptide.cpp:379-408, :570-591 for one positional isoform: at every residue step of
// the traversal (the last residue included) the m/z of each neutral-loss variant.
struct PaFragArgs {
    uint64_t sig;            // bit j = modifiable site j (N->C) carries the mod
    char type;
    int charge;              // 0 = neutral mass
    float* out_mz;           // [L][16]
    int32_t* out_nvar;       // [L]
};

__global__ void __launch_bounds__(32) k_fragment_table(PaCfg cfg, PaBatchDev b, PaFragArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PsmSmem* sm = (PsmSmem*)smem_raw;
    PsmInfo info;
    pa_setup_psm(cfg, b, 0, sm, info, false);
    if ((threadIdx.x & 31) != 0) return;
    uint64_t mlo, mhi;
    pa_sites_to_mask(sm, a.sig, mlo, mhi);
    const int L = info.L;
    const bool fwd = (a.type == 'b' || a.type == 'c');
    double a1, a2;
    pa_type_consts(a.type, a1, a2);
    float run = 0.f;
    int nls = 0;
    for (int step = 0; step < L; step++) {
        const int i = fwd ? step : L - 1 - step;
        const int st = (int)(((i < 64) ? (mlo >> i) : (mhi >> (i - 64))) & 1ull);
        const float r = sm->res[i][st];
        run = (step == 0) ? r : __fadd_rn(r, run);
        int nv = 1;
        if (cfg.has_nl) {
            int idx = sm->nlidx[i][st];
            if (idx) nls = pa_nl_bump(nls, idx);
            nv = cfg.nl_nvar[nls];
        }
        a.out_nvar[step] = nv;
        for (int v = 0; v < nv; v++) {
            float base = run;
            if (cfg.has_nl) base = __fsub_rn(run, cfg.nl_sums[nls * 16 + v]);
            const double d = __dsub_rn(__dadd_rn((double)base, a1), a2);
            a.out_mz[step * 16 + v] = pa_charge_mz(d, a.charge);
        }
    }
}

// cpp/ModifiedPeptide.cpp:259-320 (getSiteDeterminingIons) for one ion type, charges 1..max_charge
struct PaSdiArgs {
    uint64_t sig_a, sig_b;
    char type;
    int max_charge;
    float* out_a; float* out_b;      // survivors, ascending
    int32_t* counts;                 // [2]
    float* g_lists;                  // scratch for lists longer than PA_LCAP
    int64_t list_stride;
};

__global__ void __launch_bounds__(32) k_sdi_probe(PaCfg cfg, PaBatchDev b, PaSdiArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SelSmem* sm = (SelSmem*)smem_raw;
    PsmInfo info;
    pa_setup_psm(cfg, b, 0, &sm->psm, info, false);
    info.Z = a.max_charge;
    info.R = 0; info.gp = nullptr; info.pk = nullptr; info.cell = nullptr;
    info.cell_base = 0.f; info.cell_inv = 0.f;
    float* raw0 = sm->raw[0]; float* raw1 = sm->raw[1]; float* srt0 = sm->srt[0]; float* srt1 = sm->srt[1];
    long long per_type = (long long)(info.L > 1 ? info.L - 1 : 1) * cfg.nvar_cap * info.Z;
    if (per_type > PA_LCAP) {
        float* g = a.g_lists;
        raw0 = g; raw1 = g + a.list_stride; srt0 = g + 2 * a.list_stride; srt1 = g + 3 * a.list_stride;
    }
    uint64_t alo, ahi, blo, bhi;
    pa_sites_to_mask(&sm->psm, a.sig_a, alo, ahi);
    pa_sites_to_mask(&sm->psm, a.sig_b, blo, bhi);
    int hA = 0, hB = 0, tA = 0, tB = 0;
    pa_sdi_type(cfg, sm, info, a.type, alo, ahi, blo, bhi, raw0, raw1, srt0, srt1, 0, hA, tA, hB, tB);
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < tA; e += 32) a.out_a[e] = raw0[e];
    for (int e = lane; e < tB; e += 32) a.out_b[e] = raw1[e];
    if (lane == 0) { a.counts[0] = tA; a.counts[1] = tB; }
}

// cpp/Util.cpp:16-83: op 0 log_sum(x, y); 1 log_bin_coef(k, n); 2 log_pmf; 3 log_pvalue; 4 log10_pvalue
struct PaMathArgs {
    int op, n;
    const float* x; const float* y;          // op 0
    const int32_t* k; const int32_t* tr;     // ops 1-4: successes, trials
    const double* logd;                      // log((double)m), m <= max trials (host libm, as in K0)
    float lps, lpf;                          // logf(p), (float)log(1 - p)
    double log10e;
    float* out;
};

__device__ __forceinline__ float pa_log_bin_coef(const double* logd, int k, int n) {
    int kk = (n - k < k) ? n - k : k;
    float c = 0.f;
    for (int m = n - kk + 1; m <= n; m++) c = __double2float_rn(__dadd_rn((double)c, logd[m]));
    for (int m = 2; m <= kk; m++) c = __double2float_rn(__dsub_rn((double)c, logd[m]));
    return c;
}

__global__ void k_math_probe(PaMathArgs a) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.n) return;
    if (a.op == 0) { a.out[q] = pa_log_sum(a.x[q], a.y[q]); return; }
    const int k = a.k[q], n = a.tr[q];
    if (a.op == 1) { a.out[q] = pa_log_bin_coef(a.logd, k, n); return; }
    auto pmf = [&](int kk) {
        return __fadd_rn(__fadd_rn(pa_log_bin_coef(a.logd, kk, n), __fmul_rn((float)kk, a.lps)), __fmul_rn((float)(n - kk), a.lpf));
    };
    if (a.op == 2) { a.out[q] = pmf(k); return; }
    float tail = 0.f;
    if (k > 0) {
        tail = __int_as_float(0xff800000);
        for (int kk = n; kk >= k; kk--) tail = pa_log_sum(tail, pmf(kk));
    }
    a.out[q] = (a.op == 3) ? tail : __double2float_rn(__dmul_rn(a.log10e, (double)tail));
}
