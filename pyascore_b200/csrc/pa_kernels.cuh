// Kernels of libpyascore_b200 (sm_100a).  One warp owns one unit of domain work:
//   K0 k_tail_table   one block per trial count n: float32-faithful binomial score table
//   K1 k_bin_topn     one warp per spectrum: top-n_top peaks per bin_size-Th bin (BinnedSpectra)
//   P1 k_plan         one thread per PSM: validation, #sites, #isoforms, work units
//   K2 k_count_score  one warp per unit of <=1024 positional isoforms: fragments, matches, PepScore
//   K3 k_select       one warp per PSM: reference-order best isoform, Ascores, alternative sites
// Nothing here is a dense contraction: no tensor cores.  The work is integer / float32 / a little
// FP64 per fragment, bound by instruction issue and shared-memory latency.
#pragma once
#include "pa_device.cuh"

// ---------------------------------------------------------------------------------------------
// K0: score table  T[(n(n+1)/2 + k) * 10 + d] = | -10 * log10 P(X >= k) |,  X ~ Binomial(n, p_d)
// cpp/Util.cpp:28-83 + cpp/Ascore.cpp:23-36, :127-133, in the reference's float32 rounding order.
// logd[m] = log((double)m) and lps/lpf = logf(p_d), (float)log(1-p_d) are host constants from the
// platform libm (the reference gets them from the same place).
// ---------------------------------------------------------------------------------------------
struct PaTailArgs {
    float* T;
    const double* logd;
    float lps[PA_N_TOP], lpf[PA_N_TOP];
    double log10e;
};

__global__ void __launch_bounds__(128) k_tail_table(PaTailArgs a, int n0, int n1) {
    extern __shared__ float s_lbc[];
    const int n = n0 + blockIdx.x;
    if (n > n1) return;
    for (int k = threadIdx.x; k <= n; k += blockDim.x) {
        int kk = (n - k < k) ? n - k : k;
        float c = 0.f;
        for (int m = n - kk + 1; m <= n; m++) c = __double2float_rn(__dadd_rn((double)c, a.logd[m]));
        for (int m = 2; m <= kk; m++) c = __double2float_rn(__dsub_rn((double)c, a.logd[m]));
        s_lbc[k] = c;
    }
    __syncthreads();
    if (threadIdx.x < PA_N_TOP) {
        const int d = threadIdx.x;
        const float lps = a.lps[d], lpf = a.lpf[d];
        float tail = __int_as_float(0xff800000);   // -inf
        a.T[pa_tab_index(n, 0, d)] = fabsf(__fmul_rn(-10.f, __double2float_rn(__dmul_rn(a.log10e, 0.0))));
        for (int k = n; k >= 1; k--) {
            float pmf = __fadd_rn(__fadd_rn(s_lbc[k], __fmul_rn((float)k, lps)), __fmul_rn((float)(n - k), lpf));
            tail = pa_log_sum(tail, pmf);
            float l10 = __double2float_rn(__dmul_rn(a.log10e, (double)tail));
            a.T[pa_tab_index(n, k, d)] = fabsf(__fmul_rn(-10.f, l10));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1: BinnedSpectra.  cpp/Spectra.cpp:43-68 (bounds + bin index), :24-41 (top n_top by intensity).
// Fast path: the spectrum fits the warp's shared-memory slot and is m/z-sorted, so a bin is a
// contiguous run and each lane ranks its peak by scanning outwards inside the run.
// General path (unsorted or oversized spectra): all-pairs ranking through global scratch.
// Output per spectrum: retained peaks as (float)mz ascending + rank, stored at the input offsets.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pa_inten_key(double x) {
    uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct PaBinArgs {
    const int64_t* spec_off;
    const double* mz;
    const double* inten;
    int64_t peak_base;       // spec_off values are relative to this
    int64_t n_spec;
    float* rmz;
    uint8_t* rrank;
    int32_t* rcount;
    uint8_t* ctab;           // per spectrum PA_NCELL bytes (m/z cell index of the retained peaks)
    float2* chead;           // per spectrum {cell base, 1 / cell width}
    int32_t* g_bin;          // scratch, one per peak
    uint8_t* g_tmp;          // scratch, one per peak
    float bin_size;
    int n_top;
    int cap;                 // peaks per warp slot in shared memory
};

#define PA_NBIN_SMEM 128     // bins whose [start,end) ranges are tabulated in shared memory

// shared memory per warp slot: mz f64[cap] (later {float mz, int bin}) | key u64[cap] |
// bstart u16[128] | bend u16[128] | cell u8[256]
#define PA_BIN_SLOT_BYTES(cap) ((size_t)(cap) * 16 + PA_NBIN_SMEM * 4 + PA_NCELL)

__device__ __forceinline__ void pa_cp_async8(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(256) k_bin_topn(PaBinArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int cap = a.cap;
    unsigned char* slot = smem_raw + (size_t)wib * PA_BIN_SLOT_BYTES(cap);
    double* s_mz = (double*)slot;
    float2* s_mb = (float2*)slot;                 // after binning: {float mz, bin as int bits}, in place
    float* s_out = (float*)slot;                  // retained (float)mz, compacted in place
    uint64_t* s_key = (uint64_t*)(slot + (size_t)cap * 8);
    uint16_t* s_bstart = (uint16_t*)(slot + (size_t)cap * 16);
    uint16_t* s_bend = s_bstart + PA_NBIN_SMEM;
    uint8_t* s_cell = (uint8_t*)(s_bend + PA_NBIN_SMEM);
    const int64_t gw = (int64_t)blockIdx.x * wpb + wib, nw = (int64_t)gridDim.x * wpb;
    const int n_top = a.n_top;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);

    for (int64_t s = gw; s < a.n_spec; s += nw) {
        const int64_t off = a.spec_off[s] - a.peak_base;
        const int P = (int)(a.spec_off[s + 1] - a.spec_off[s]);
        if (P <= 0) { if (lane == 0) { a.rcount[s] = 0; a.chead[s] = make_float2(0.f, 0.f); } continue; }
        const bool fits = P <= cap;
        double mn = INF, mx = -INF;
        bool sorted = true;
        if (fits) {
            // stage the whole spectrum with asynchronous 8-byte copies: every load of the warp is
            // in flight at once (the arrays are only 8-byte aligned at a CSR offset)
            for (int i = lane; i < P; i += 32) {
                pa_cp_async8(&s_mz[i], a.mz + off + i);
                pa_cp_async8(&s_key[i], a.inten + off + i);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncwarp();
            for (int i = lane; i < P; i += 32) {
                const double m = s_mz[i];
                const double prev = i > 0 ? s_mz[i - 1] : -INF;
                if (m < prev) sorted = false;
                mn = m < mn ? m : mn;
                mx = m > mx ? m : mx;
            }
        } else {
            double carry = -INF;
            for (int base = 0; base < P; base += 32) {
                int i = base + lane;
                double m = 0.;
                if (i < P) m = a.mz[off + i];
                double prev = __shfl_up_sync(PA_FULL, m, 1);
                if (lane == 0) prev = carry;
                if (i < P) {
                    if (m < prev) sorted = false;
                    mn = m < mn ? m : mn;
                    mx = m > mx ? m : mx;
                }
                carry = __shfl_sync(PA_FULL, m, 31);
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            double t = __shfl_xor_sync(PA_FULL, mn, o); mn = t < mn ? t : mn;
            t = __shfl_xor_sync(PA_FULL, mx, o); mx = t > mx ? t : mx;
        }
        sorted = __all_sync(PA_FULL, sorted);
        // cpp/Spectra.cpp:46-48: the 100 is a literal there, independent of bin_size
        const float min_mz = __double2float_rn(__dmul_rn(floor(__ddiv_rn(mn, 100.)), 100.));
        const float max_mz = __double2float_rn(__dmul_rn(ceil(__ddiv_rn(mx, 100.)), 100.));
        long long n_bins = (long long)ceilf(__fdiv_rn(__fsub_rn(max_mz, min_mz), a.bin_size));
        if (n_bins < 1) n_bins = 1;     // degenerate spectrum: undefined in the reference
        const double dmin = (double)min_mz, dbs = (double)a.bin_size;

        if (fits && sorted) {
            const bool tab = n_bins <= PA_NBIN_SMEM;
            // bins; a sorted spectrum makes every bin one contiguous run [bstart, bend)
            int carry_bin = -1;
            for (int base = 0; base < P; base += 32) {
                const int i = base + lane;
                int bq = -1;
                if (i < P) {
                    const double m = s_mz[i];
                    double q = floor(__ddiv_rn(__dsub_rn(m, dmin), dbs));
                    long long b64 = (long long)q;
                    if (b64 > n_bins - 1) b64 = n_bins - 1;
                    bq = (int)b64;
                    s_mb[i] = make_float2(__double2float_rn(m), __int_as_float(bq));   // own slot, in place
                    s_key[i] = pa_inten_key(__longlong_as_double((long long)s_key[i]));
                }
                int bprev = __shfl_up_sync(PA_FULL, bq, 1);
                if (lane == 0) bprev = carry_bin;
                if (tab && i < P) {
                    if (bq != bprev) { s_bstart[bq] = (uint16_t)i; if (i > 0) s_bend[bprev] = (uint16_t)i; }
                    if (i == P - 1) s_bend[bq] = (uint16_t)P;
                }
                carry_bin = __shfl_sync(PA_FULL, bq, 31);
            }
            __syncwarp();
            int out = 0;
            for (int base = 0; base < P; base += 32) {
                const int i = base + lane;
                int cnt = n_top;
                float mzf = 0.f;
                if (i < P) {
                    const float2 mb = s_mb[i];
                    mzf = mb.x;
                    const int bq = __float_as_int(mb.y);
                    const uint64_t ki = s_key[i];
                    cnt = 0;
                    if (tab) {
                        // rank = peaks of the same bin that beat this one; an equal intensity wins
                        // only from an earlier index: (kj + [j < i]) > ki covers both sides of i
                        const int b0 = s_bstart[bq], b1 = s_bend[bq];
                        int j = b0;
                        for (; j + 4 <= b1; j += 4) {
                            const uint64_t k0 = s_key[j], k1 = s_key[j + 1], k2 = s_key[j + 2], k3 = s_key[j + 3];
                            cnt += ((k0 + (uint64_t)(j < i)) > ki) + ((k1 + (uint64_t)(j + 1 < i)) > ki) +
                                   ((k2 + (uint64_t)(j + 2 < i)) > ki) + ((k3 + (uint64_t)(j + 3 < i)) > ki);
                        }
                        for (; j < b1; j++) cnt += ((s_key[j] + (uint64_t)(j < i)) > ki);
                    } else {
                        for (int j = i - 1; j >= 0 && cnt < n_top && __float_as_int(s_mb[j].y) == bq; j--) cnt += (s_key[j] >= ki);
                        for (int j = i + 1; j < P && cnt < n_top && __float_as_int(s_mb[j].y) == bq; j++) cnt += (s_key[j] > ki);
                    }
                }
                const bool keep = cnt < n_top;
                unsigned bal = __ballot_sync(PA_FULL, keep);   // every lane has read its own s_mb slot by now
                if (keep) {
                    int pos = out + __popc(bal & ((1u << lane) - 1u));
                    a.rmz[off + pos] = mzf;
                    a.rrank[off + pos] = (uint8_t)cnt;
                    if (tab) s_out[pos] = mzf;      // pos <= i: only slots this or earlier chunks own
                }
                out += __popc(bal);
                __syncwarp();
            }
            if (lane == 0) a.rcount[s] = out;
            // m/z cell index over the retained peaks (consumers: pa_match_rank)
            if (tab && out <= PA_RCAP && out > 0) {
                __syncwarp();
                const float cbase = s_out[0];
                const float cinv = pa_cell_inv(cbase, s_out[out - 1]);
                for (int j = lane; j < out; j += 32) {
                    const int cj = pa_cell(s_out[j], cbase, cinv);
                    const int cp = j > 0 ? pa_cell(s_out[j - 1], cbase, cinv) : -1;
                    for (int c = cp + 1; c <= cj; c++) s_cell[c] = (uint8_t)j;
                    if (j == out - 1) for (int c = cj + 1; c < PA_NCELL; c++) s_cell[c] = (uint8_t)out;
                }
                __syncwarp();
                ((unsigned long long*)(a.ctab + (size_t)s * PA_NCELL))[lane] = ((const unsigned long long*)s_cell)[lane];
                if (lane == 0) a.chead[s] = make_float2(cbase, cinv);
            } else if (lane == 0) a.chead[s] = make_float2(0.f, 0.f);
        } else {
            // general path through global scratch
            for (int i = lane; i < P; i += 32) {
                double q = floor(__ddiv_rn(__dsub_rn(a.mz[off + i], dmin), dbs));
                long long bq = (long long)q;
                if (bq > n_bins - 1) bq = n_bins - 1;
                a.g_bin[off + i] = (int32_t)bq;
            }
            __syncwarp();
            for (int i = lane; i < P; i += 32) {
                const int bq = a.g_bin[off + i];
                const uint64_t ki = pa_inten_key(a.inten[off + i]);
                int cnt = 0;
                for (int j = 0; j < P && cnt < n_top; j++) {
                    if (a.g_bin[off + j] != bq || j == i) continue;
                    uint64_t kj = pa_inten_key(a.inten[off + j]);
                    cnt += (kj > ki) || (kj == ki && j < i);
                }
                a.g_tmp[off + i] = (uint8_t)(cnt < n_top ? cnt : 255);
            }
            __syncwarp();
            int total = 0;
            for (int base = 0; base < P; base += 32) {
                int i = base + lane;
                bool keep = (i < P) && a.g_tmp[off + i] != 255;
                if (keep) {
                    const float fi = __double2float_rn(a.mz[off + i]);
                    int pos = 0;
                    for (int j = 0; j < P; j++) {
                        if (a.g_tmp[off + j] == 255 || j == i) continue;
                        float fj = __double2float_rn(a.mz[off + j]);
                        pos += (fj < fi) || (fj == fi && j < i);
                    }
                    a.rmz[off + pos] = fi;
                    a.rrank[off + pos] = a.g_tmp[off + i];
                }
                total += __popc(__ballot_sync(PA_FULL, keep));
            }
            if (lane == 0) { a.rcount[s] = total; a.chead[s] = make_float2(0.f, 0.f); }
        }
        __syncwarp();
    }
}

__global__ void k_max_peaks(const int64_t* spec_off, int64_t n_spec, int* out) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int v = 0;
    if (q < n_spec) { int64_t d = spec_off[q + 1] - spec_off[q]; v = d > (1 << 20) ? (1 << 20) : (int)d; }
    for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(PA_FULL, v, o); v = t > v ? t : v; }
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

// ---------------------------------------------------------------------------------------------
// P1: per-PSM planning (validation, sites, isoform count, K2 work units)
// ---------------------------------------------------------------------------------------------
struct PaPlanOut {
    int32_t* psm_S;
    int32_t* psm_status;
    int64_t* psm_I;          // [n_psm+1], last = 0 (input of the exclusive scan)
    int32_t* psm_units;      // [n_psm+1], last = 0
    unsigned long long* combo_bits;  // [64]: bit k of row S set when (S,k) occurs with > 1 isoform
    int* max_frag;           // max fragments per isoform (all types/charges) over the chunk
    int* max_list;           // max fragments per (isoform, type) over the chunk (K3 list size)
};

__global__ void __launch_bounds__(256) k_plan(PaCfg cfg, PaBatchDev b, int64_t n_psm, PaPlanOut o) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_psm) return;
    const int off = b.pep_off[p];
    const int L = b.pep_off[p + 1] - off;
    const int k = b.n_mod[p], Z = b.max_charge[p];
    const int sp = b.psm_spec[p];
    int status = PA_PSM_OK, S = 0;
    if (sp < 0 || sp >= b.n_spec || k < 0 || Z < 1 || L < 1) status = PA_PSM_BAD_INDEX;  // Z == 0 crashes the reference
    else if (L > PA_MAX_PEPTIDE) status = PA_PSM_TOO_LONG;
    else {
        for (int i = 0; i < L; i++) {
            int c = (int)b.pep[off + i] - 'A';
            if (c < 0 || c >= 26 || isnan(cfg.res_mass[c])) { status = PA_PSM_BAD_RESIDUE; break; }
            bool site = ((cfg.mod_letters >> c) & 1u) || (cfg.allow_n && i == 0) || (cfg.allow_c && i == L - 1);
            S += site;
        }
        if (status == PA_PSM_OK && b.aux_off != nullptr)
            for (int a = b.aux_off[p]; a < b.aux_off[p + 1]; a++)
                if (b.aux_pos[a] > (uint32_t)L) status = PA_PSM_BAD_AUX;
        if (status == PA_PSM_OK && S > PA_MAX_SITES) status = PA_PSM_TOO_MANY_SITES;
        if (status == PA_PSM_OK && b.rcount[sp] <= 0) status = PA_PSM_EMPTY_SPECTRUM;
    }
    int64_t I = 0;
    if (status == PA_PSM_OK) {
        long long per_type = (long long)(L > 1 ? L - 1 : 1) * cfg.nvar_cap * Z;
        long long nf = per_type * cfg.n_types;
        if (nf > PA_MAX_FRAGMENTS) status = PA_PSM_TOO_MANY_FRAGMENTS;
        else {
            uint32_t c = (k <= S) ? cfg.binom[S * 64 + k] : 0u;
            if ((long long)c > PA_MAX_ISOFORMS) status = PA_PSM_TOO_MANY_ISOFORMS;
            else {
                I = c;
                atomicMax(o.max_frag, (int)nf);
                atomicMax(o.max_list, (int)per_type);
                if (I > 1) atomicOr(&o.combo_bits[S], 1ull << k);
            }
        }
    }
    o.psm_S[p] = S;
    o.psm_status[p] = status;
    o.psm_I[p] = I;
    o.psm_units[p] = (int32_t)((I + PA_UNIT - 1) / PA_UNIT);
}

__global__ void k_expand_units(int64_t n_psm, const int32_t* unit_off, int32_t* unit_psm) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_psm) return;
    for (int u = unit_off[p]; u < unit_off[p + 1]; u++) unit_psm[u] = (int32_t)p;
}

// ---------------------------------------------------------------------------------------------
// K2: per-isoform fragment generation, peak matching, per-depth counts and PepScore.
// cpp/ModifiedPeptide.cpp:379-408, :500-524, :570-591 (fragments), :126-150 (matching),
// cpp/Ascore.cpp:53-121 (counts), :123-139 (scores).  One lane owns one isoform at a time and
// performs the reference's sequential float32 running sum for it, so every fragment m/z has the
// reference's bits.
// ---------------------------------------------------------------------------------------------
struct PaIso {                 // per-isoform records of a chunk (SoA), indexed iso_off[p] + idx
    unsigned long long* lo;    // cumulative counts depth 0..4, 12 bits each
    unsigned long long* hi;    // cumulative counts depth 5..9
    uint32_t* nfrag;           // total_fragments
    float* w;                  // weighted PepScore
};

struct PaCountArgs {
    int64_t n_units;
    const int32_t* unit_psm;
    const int32_t* unit_off;
    const int64_t* iso_off;
    const int32_t* psm_S;
    const int32_t* psm_status;
    PaIso iso;
    unsigned long long* n_lookups;   // counter
};

// Walk the fragments of ion types [t0, t1) of the isoform with residue mask (mlo,mhi); returns
// packed non-cumulative per-rank counts.
template <bool HAS_NL>
__device__ __forceinline__ void pa_walk_isoform(const PaCfg& cfg, const PsmSmem* sm, const PsmInfo& info,
                                                uint64_t mlo, uint64_t mhi, int t0, int t1,
                                                unsigned long long& clo, unsigned long long& chi,
                                                uint32_t& nfrag) {
    const int L = info.L, Z = info.Z;
    clo = 0; chi = 0; nfrag = 0;
    // L == 1: the walk starts on the last residue and the reference's end test lets all but
    // the last neutral-loss variant through (cpp/ModifiedPeptide.cpp:516-524)
    const int steps = (L == 1) ? 1 : L - 1;
    for (int t = t0; t < t1; t++) {
        const char type = cfg.types[t];
        const bool fwd = (type == 'b' || type == 'c');
        double a1, a2;
        pa_type_consts(type, a1, a2);
        float run = 0.f;
        int nls = 0;
        for (int step = 0; step < steps; step++) {
            const int i = fwd ? step : L - 1 - step;
            const int st = (int)(((i < 64) ? (mlo >> i) : (mhi >> (i - 64))) & 1ull);
            const float r = sm->res[i][st];
            run = (step == 0) ? r : __fadd_rn(r, run);
            int nv = 1;
            if (HAS_NL) {
                int idx = sm->nlidx[i][st];
                if (idx) nls = pa_nl_bump(nls, idx);
                nv = cfg.nl_nvar[nls];
            }
            if (L == 1) nv -= 1;
            for (int v = 0; v < nv; v++) {
                float base = run;
                if (HAS_NL) base = __fsub_rn(run, __ldg(cfg.nl_sums + nls * 16 + v));
                const double d = __dsub_rn(__dadd_rn((double)base, a1), a2);
                for (int z = 1; z <= Z; z++) {
                    const float f = pa_charge_mz(d, z);
                    const int rk = pa_match_rank(info, f, cfg.err, cfg.err_gt_half);
                    if (rk < 5) clo += 1ull << (12 * rk);
                    else if (rk < 10) chi += 1ull << (12 * (rk - 5));
                }
                nfrag += Z;
            }
        }
    }
}

// packed per-rank counts -> packed cumulative counts (cpp/Ascore.cpp:113-117)
__device__ __forceinline__ void pa_cumulate(unsigned long long& lo, unsigned long long& hi) {
    // prefix sums of 12-bit fields; totals stay < 4096 (PA_MAX_FRAGMENTS)
    lo += lo << 12; lo += lo << 24; lo += lo << 48;        // fields 0..4 (60 bits): inclusive scan
    unsigned long long top = (lo >> 48) & 0xfffull;        // cumulative count of ranks 0..4
    hi += hi << 12; hi += hi << 24; hi += hi << 48;
    hi += top * 0x001001001001001ull;
    lo &= 0x0fffffffffffffffull; hi &= 0x0fffffffffffffffull;
}

__device__ __forceinline__ int pa_cum_get(unsigned long long lo, unsigned long long hi, int d) {
    return (int)(((d < 5) ? (lo >> (12 * d)) : (hi >> (12 * (d - 5)))) & 0xfffull);
}

// cpp/Ascore.cpp:123-139
__device__ __forceinline__ float pa_weighted(const PaCfg& cfg, unsigned long long lo, unsigned long long hi, int n) {
    double acc = 0.;
#pragma unroll
    for (int d = 0; d < PA_N_TOP; d++) {
        float sc = __ldg(cfg.T + pa_tab_index(n, pa_cum_get(lo, hi, d), d));
        acc = __dadd_rn(acc, (double)__fmul_rn(cfg.weights[d], sc));
    }
    return __double2float_rn(acc);
}

__device__ __forceinline__ void pa_sites_to_mask(const PsmSmem* sm, uint64_t bits, uint64_t& mlo, uint64_t& mhi) {
    mlo = 0; mhi = 0;
    while (bits) {
        int j = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        int pos = sm->site_pos[j];
        if (pos < 64) mlo |= 1ull << pos; else mhi |= 1ull << (pos - 64);
    }
}

// PAIR: exactly two ion types (the default "by"): adjacent lanes take the two types of one isoform
// and add their counts with one shuffle, so small PSMs keep twice as many lanes busy.
template <bool HAS_NL, bool PAIR>
__global__ void __launch_bounds__(256) k_count_score(PaCfg cfg, PaBatchDev b, PaCountArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    PsmSmem* sm = (PsmSmem*)smem_raw + wib;
    const int64_t gw = (int64_t)blockIdx.x * wpb + wib, nw = (int64_t)gridDim.x * wpb;
    int64_t cur = -1;
    PsmInfo info;
    unsigned long long lookups = 0;
    for (int64_t u = gw; u < a.n_units; u += nw) {
        const int64_t p = a.unit_psm[u];
        if (p != cur) { pa_setup_psm(cfg, b, p, sm, info, true); cur = p; }
        const int S = a.psm_S[p], k = info.k;
        const int64_t I = a.iso_off[p + 1] - a.iso_off[p];
        const int64_t first = (int64_t)(u - a.unit_off[p]) * PA_UNIT;
        const int64_t cnt = (I - first < PA_UNIT) ? I - first : PA_UNIT;
        const int64_t items = PAIR ? 2 * cnt : cnt;
        for (int64_t q0 = 0; q0 < items; q0 += 32) {
            const int64_t q = q0 + lane;
            const bool active = q < items;
            const uint32_t idx = (uint32_t)(first + (PAIR ? (q >> 1) : q));
            unsigned long long clo = 0, chi = 0; uint32_t nf = 0;
            if (active) {
                const uint64_t bits = pa_unrank(cfg.binom, S, k, idx);
                uint64_t mlo, mhi;
                pa_sites_to_mask(sm, bits, mlo, mhi);
                const int t0 = PAIR ? (int)(q & 1) : 0, t1 = PAIR ? t0 + 1 : cfg.n_types;
                pa_walk_isoform<HAS_NL>(cfg, sm, info, mlo, mhi, t0, t1, clo, chi, nf);
                lookups += nf;
            }
            if (PAIR) {
                clo += __shfl_xor_sync(PA_FULL, clo, 1);
                chi += __shfl_xor_sync(PA_FULL, chi, 1);
                nf += __shfl_xor_sync(PA_FULL, nf, 1);
            }
            if (active && (!PAIR || (lane & 1) == 0)) {
                pa_cumulate(clo, chi);
                const int64_t g = a.iso_off[p] + idx;
                a.iso.lo[g] = clo; a.iso.hi[g] = chi; a.iso.nfrag[g] = nf;
                a.iso.w[g] = pa_weighted(cfg, clo, chi, (int)nf);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) lookups += __shfl_xor_sync(PA_FULL, lookups, o);
    if (lane == 0 && lookups) atomicAdd(a.n_lookups, lookups);
}

// ---------------------------------------------------------------------------------------------
// K3: best isoform in the reference's order, Ascores and alternative sites.
// cpp/Ascore.cpp:38-51 (unambiguous), :141-146 (std::sort), :212-254 (calculateAscores),
// :157-210 (calculateAmbiguity), cpp/ModifiedPeptide.cpp:259-320 (site-determining ions).
// ---------------------------------------------------------------------------------------------
#define PA_LCAP 256      // fragments per (isoform, ion type) list staged in shared memory (power of two)
#define PA_SORTCAP 256   // isoforms sortable in shared memory (aliases the list area)

struct SelSmem {
    PsmSmem psm;
    float run[2][PA_LMAX];          // running float sums of the two isoforms
    uint8_t nls[2][PA_LMAX];        // neutral-loss state after each step
    uint16_t foff[2][PA_LMAX + 4];  // fragment offset of each step (in units of one charge)
    alignas(8) float raw[2][PA_LCAP];  // unsorted fragment lists A, B (also the std::sort arena)
    float srt[2][PA_LCAP];          // sorted
};

struct PaSelArgs {
    int64_t n_psm;
    const int64_t* iso_off;
    const int32_t* psm_S;
    const int32_t* psm_status;
    PaIso iso;
    const int64_t* perm_off;     // [64*64] offset into perm_pool of the hash-order list of (S,k), -1 = none
    const uint32_t* perm_pool;   // isoform index (lexicographic rank) at each hash-iteration position
    const int64_t* mod_off;
    // outputs (device)
    uint64_t* best_sig;
    float* best_score;
    int64_t* n_iso;
    int32_t* n_sites;
    float* ascores;
    uint64_t* alt_sites;
    int32_t* psm_status_out;
    // scratch
    unsigned long long* g_sort;  // per isoform (iso_off), for std::sort emulation beyond PA_SORTCAP
    int64_t mod_lo;              // absolute index of the chunk's first mod entry
    uint32_t* best_idx;          // [n_psm] best isoform (lexicographic rank), 0xffffffff = none
    int32_t* mod_psm;            // [entries] chunk-relative PSM of each mod entry, -1 = no Ascore to compute
    unsigned long long* tie;     // [entries] tied best competitors of each mod entry
};

// --- libstdc++ std::sort (introsort + final insertion sort), comparator a.w > b.w ------------
// bits/stl_algo.h of GCC 13, as in SURVEY.md appendix A.2.  Elements are (float w, uint32 id)
// packed in 64 bits: w in the high word.  Run by ONE lane.
__device__ __forceinline__ float srt_w(unsigned long long e) { return __int_as_float((int)(e >> 32)); }
#define SRT_CMP(x, y) (srt_w(x) > srt_w(y))

__device__ void srt_adjust_heap(unsigned long long* a, long hole, long len, unsigned long long v) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (SRT_CMP(a[child], a[child - 1])) child--;
        a[hole] = a[child]; hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[hole] = a[child - 1]; hole = child - 1;
    }
    long parent = (hole - 1) / 2;
    while (hole > top && SRT_CMP(a[parent], v)) { a[hole] = a[parent]; hole = parent; parent = (hole - 1) / 2; }
    a[hole] = v;
}

__device__ void srt_heap_sort(unsigned long long* a, long n) {
    if (n < 2) return;
    for (long parent = (n - 2) / 2;; parent--) {
        srt_adjust_heap(a, parent, n, a[parent]);
        if (parent == 0) break;
    }
    for (long last = n; last > 1;) {
        --last;
        unsigned long long v = a[last];
        a[last] = a[0];
        srt_adjust_heap(a, 0, last, v);
    }
}

__device__ __forceinline__ void srt_linear_insert(unsigned long long* a, long i) {
    unsigned long long v = a[i];
    long j = i - 1;
    while (SRT_CMP(v, a[j])) { a[j + 1] = a[j]; j--; }
    a[j + 1] = v;
}

__device__ __forceinline__ void srt_insertion(unsigned long long* a, long first, long last) {
    for (long i = first + 1; i < last; i++) {
        if (SRT_CMP(a[i], a[first])) {
            unsigned long long v = a[i];
            for (long j = i; j > first; j--) a[j] = a[j - 1];
            a[first] = v;
        } else srt_linear_insert(a, i);
    }
}

__device__ void pa_gcc_sort(unsigned long long* a, long n) {
    if (n < 2) return;
    long lg = 0;
    for (long t = n; t > 1; t >>= 1) lg++;
    // __introsort_loop with an explicit stack instead of recursion (the recursion is on the right part)
    long stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        sp--;
        long first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        // the reference processes [cut,last) recursively FIRST, then loops on [first,cut).  The
        // order in which disjoint ranges are partitioned does not change the result.
        while (last - first > 16) {
            if (depth == 0) { srt_heap_sort(a + first, last - first); break; }
            --depth;
            long mid = first + (last - first) / 2;
            long x = first + 1, y = mid, z = last - 1, pick;
            if (SRT_CMP(a[x], a[y])) { if (SRT_CMP(a[y], a[z])) pick = y; else if (SRT_CMP(a[x], a[z])) pick = z; else pick = x; }
            else if (SRT_CMP(a[x], a[z])) pick = x; else if (SRT_CMP(a[y], a[z])) pick = z; else pick = y;
            { unsigned long long t = a[first]; a[first] = a[pick]; a[pick] = t; }
            long f = first + 1, l = last;
            for (;;) {
                while (SRT_CMP(a[f], a[first])) f++;
                --l;
                while (SRT_CMP(a[first], a[l])) l--;
                if (!(f < l)) break;
                { unsigned long long t = a[f]; a[f] = a[l]; a[l] = t; }
                f++;
            }
            if (sp < 64) { stk_first[sp] = f; stk_last[sp] = last; stk_depth[sp] = depth; sp++; }
            last = f;
        }
    }
    if (n > 16) { srt_insertion(a, 0, 16); for (long i = 16; i < n; i++) srt_linear_insert(a, i); }
    else srt_insertion(a, 0, n);
}

// Only the element std::sort would leave at index 0 is needed on the device (the full listing
// order is produced on the host by pa_fetch_pep_scores).  Introsort never moves an element from
// the right part of a partition into the left part afterwards, and the closing insertion sort moves
// an element left only past strictly smaller scores, so position 0 is decided by the chain of
// LEFTMOST partitions alone: replay those partitions (O(n) instead of O(n log n)), then take the
// first maximal element of the final left segment (the insertion sort is stable there).
__device__ uint32_t pa_gcc_sort_front(unsigned long long* a, long n) {
    if (n < 1) return 0xffffffffu;
    long lg = 0;
    for (long t = n; t > 1; t >>= 1) lg++;
    long last = n, depth = 2 * lg;
    const long first = 0;
    while (last - first > 16) {
        if (depth == 0) { srt_heap_sort(a, last); return (uint32_t)(a[0] & 0xffffffffull); }
        --depth;
        long mid = first + (last - first) / 2;
        long x = first + 1, y = mid, z = last - 1, pick;
        if (SRT_CMP(a[x], a[y])) { if (SRT_CMP(a[y], a[z])) pick = y; else if (SRT_CMP(a[x], a[z])) pick = z; else pick = x; }
        else if (SRT_CMP(a[x], a[z])) pick = x; else if (SRT_CMP(a[y], a[z])) pick = z; else pick = y;
        { unsigned long long t = a[first]; a[first] = a[pick]; a[pick] = t; }
        long f = first + 1, l = last;
        const unsigned long long pv = a[first];
        for (;;) {
            while (SRT_CMP(a[f], pv)) f++;
            --l;
            while (SRT_CMP(pv, a[l])) l--;
            if (!(f < l)) break;
            { unsigned long long t = a[f]; a[f] = a[l]; a[l] = t; }
            f++;
        }
        last = f;
    }
    long bi = 0;
    for (long i = 1; i < last; i++) if (SRT_CMP(a[i], a[bi])) bi = i;
    return (uint32_t)(a[bi] & 0xffffffffull);
}

// --- site-determining ions of isoforms A (slot 0) and B (slot 1) for one ion type ------------
// Returns via hits/trials accumulators (lane-uniform).  `la`,`lb`: list pointers (smem or global).
__device__ __forceinline__ void pa_sdi_type(const PaCfg& cfg, SelSmem* sm, const PsmInfo& info, char type,
                                            uint64_t maskA_lo, uint64_t maskA_hi, uint64_t maskB_lo,
                                            uint64_t maskB_hi, float* raw0, float* raw1, float* srt0,
                                            float* srt1, int depth, int& hitsA, int& trialsA, int& hitsB,
                                            int& trialsB) {
    const int lane = threadIdx.x & 31;
    const int L = info.L, Z = info.Z;
    const bool fwd = (type == 'b' || type == 'c');
    const int steps = (L == 1) ? 1 : L - 1;     // see pa_walk_isoform for the one-residue rule
    // 1. sequential running sums (lane 0: A, lane 1: B)
    if (lane < 2) {
        const uint64_t mlo = lane ? maskB_lo : maskA_lo, mhi = lane ? maskB_hi : maskA_hi;
        float run = 0.f;
        int nls = 0, off = 0;
        for (int step = 0; step < steps; step++) {
            const int i = fwd ? step : L - 1 - step;
            const int st = (int)(((i < 64) ? (mlo >> i) : (mhi >> (i - 64))) & 1ull);
            const float r = sm->psm.res[i][st];
            run = (step == 0) ? r : __fadd_rn(r, run);
            int nv = 1;
            if (cfg.has_nl) {
                int idx = sm->psm.nlidx[i][st];
                if (idx) nls = pa_nl_bump(nls, idx);
                nv = cfg.nl_nvar[nls];
            }
            if (L == 1) nv -= 1;
            sm->run[lane][step] = run;
            sm->nls[lane][step] = (uint8_t)nls;
            sm->foff[lane][step] = (uint16_t)off;
            off += nv;
        }
        sm->foff[lane][steps] = (uint16_t)off;
    }
    __syncwarp();
    const int nA = sm->foff[0][steps] * Z, nB = sm->foff[1][steps] * Z;
    // 2. all fragments (charges 1..Z) straight into the sort arrays, padded with +inf to a power of two
    const float PINF = __int_as_float(0x7f800000);
    int NA = 32, NB = 32;
    while (NA < nA) NA <<= 1;
    while (NB < nB) NB <<= 1;
    double a1, a2;
    pa_type_consts(type, a1, a2);
    for (int w = 0; w < 2; w++) {
        float* dst = w ? srt1 : srt0;
        const int n = w ? nB : nA, N = w ? NB : NA;
        for (int e = n + lane; e < N; e += 32) dst[e] = PINF;
        for (int step = lane; step < steps; step += 32) {
            const float run = sm->run[w][step];
            const int nls = sm->nls[w][step];
            const int nv = sm->foff[w][step + 1] - sm->foff[w][step];
            const int o = sm->foff[w][step] * Z;
            for (int v = 0; v < nv; v++) {
                float base = cfg.has_nl ? __fsub_rn(run, __ldg(cfg.nl_sums + nls * 16 + v)) : run;
                const double d = __dsub_rn(__dadd_rn((double)base, a1), a2);
                for (int z = 1; z <= Z; z++) dst[o + v * Z + (z - 1)] = pa_charge_mz(d, z);
            }
        }
    }
    __syncwarp();
    // 3. sort both lists ascending: warp-wide bitonic network over shared (or global) memory.
    //    Only the sorted VALUES matter downstream, so any correct sort reproduces std::sort here.
    for (int w = 0; w < 2; w++) {
        float* arr = w ? srt1 : srt0;
        const int N = w ? NB : NA;
        for (int kk = 2; kk <= N; kk <<= 1) {
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int t = lane; t < (N >> 1); t += 32) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int l = i | j;
                    const bool asc = (i & kk) == 0;
                    const float x = arr[i], y = arr[l];
                    if ((x > y) == asc) { arr[i] = y; arr[l] = x; }
                }
                __syncwarp();
            }
        }
    }
    // 4. greedy tolerance merge (cpp/ModifiedPeptide.cpp:288-316): sequential by nature, run by
    //    lane 0 until one list is exhausted; the tail of the other list survives wholesale.
    //    Survivors are compacted at the front of raw0 / raw1.
    int cA = 0, cB = 0, ti = 0, tj = 0;
    if (lane == 0) {
        int i = 0, j = 0;
        const float err = cfg.err;
        float x = nA > 0 ? srt0[0] : PINF, y = nB > 0 ? srt1[0] : PINF;
        while (i < nA && j < nB) {
            if (fabsf(__fsub_rn(x, y)) < err) {
                i++; j++;
                x = i < nA ? srt0[i] : PINF;
                y = j < nB ? srt1[j] : PINF;
            } else if (x < y) {
                raw0[cA++] = x; i++;
                x = i < nA ? srt0[i] : PINF;
            } else {
                raw1[cB++] = y; j++;
                y = j < nB ? srt1[j] : PINF;
            }
        }
        ti = i; tj = j;
    }
    cA = __shfl_sync(PA_FULL, cA, 0);
    cB = __shfl_sync(PA_FULL, cB, 0);
    ti = __shfl_sync(PA_FULL, ti, 0);
    tj = __shfl_sync(PA_FULL, tj, 0);
    for (int e = ti + lane; e < nA; e += 32) raw0[cA + (e - ti)] = srt0[e];
    for (int e = tj + lane; e < nB; e += 32) raw1[cB + (e - tj)] = srt1[e];
    cA += nA - ti;
    cB += nB - tj;
    __syncwarp();
    // 5. hits: survivors whose matched rank <= depth
    int hA = 0, hB = 0;
    for (int e = lane; e < cA; e += 32)
        hA += pa_match_rank(info, raw0[e], cfg.err, cfg.err_gt_half) <= depth;
    for (int e = lane; e < cB; e += 32)
        hB += pa_match_rank(info, raw1[e], cfg.err, cfg.err_gt_half) <= depth;
    for (int o = 16; o > 0; o >>= 1) { hA += __shfl_xor_sync(PA_FULL, hA, o); hB += __shfl_xor_sync(PA_FULL, hB, o); }
    hitsA += hA; hitsB += hB; trialsA += cA; trialsB += cB;
    __syncwarp();
}

// cpp/Ascore.cpp:157-210.  scA/scB: the ten depth scores of the two isoforms (lane-uniform arrays).
__device__ __forceinline__ float pa_ambiguity(const PaCfg& cfg, SelSmem* sm, const PsmInfo& info, uint64_t bitsA,
                                              const float* scA, float wA, uint64_t bitsB, const float* scB,
                                              float wB, float* raw0, float* raw1, float* srt0, float* srt1) {
    if ((double)fabsf(__fsub_rn(wA, wB)) < 1e-6) return 0.f;
    float max_diff = 0.f;
    int depth = 0;
#pragma unroll
    for (int d = 0; d < PA_N_TOP; d++) {
        float diff = __fsub_rn(scA[d], scB[d]);
        if (diff > max_diff) { max_diff = diff; depth = d; }
    }
    uint64_t alo, ahi, blo, bhi;
    pa_sites_to_mask(&sm->psm, bitsA, alo, ahi);
    pa_sites_to_mask(&sm->psm, bitsB, blo, bhi);
    int hitsA = 0, hitsB = 0, trialsA = 0, trialsB = 0;
    for (int t = 0; t < cfg.n_types; t++)
        pa_sdi_type(cfg, sm, info, cfg.types[t], alo, ahi, blo, bhi, raw0, raw1, srt0, srt1, depth, hitsA,
                    trialsA, hitsB, trialsB);
    const float sA = __ldg(cfg.T + pa_tab_index(trialsA, hitsA, depth));
    const float sB = __ldg(cfg.T + pa_tab_index(trialsB, hitsB, depth));
    return __fsub_rn(sA, sB);
}

__device__ __forceinline__ void pa_depth_scores(const PaCfg& cfg, unsigned long long lo, unsigned long long hi,
                                                int n, float* sc) {
#pragma unroll
    for (int d = 0; d < PA_N_TOP; d++) sc[d] = __ldg(cfg.T + pa_tab_index(n, pa_cum_get(lo, hi, d), d));
}

// K3a: warp per PSM.  Best isoform in the reference's order + per modified site the set of tied
// best competitors (= alternative sites).  The Ascore of every (PSM, site) entry is then computed
// by k_ascore (thread per entry) or, for the rare shapes that kernel does not cover, k_ascore_generic.
__global__ void __launch_bounds__(256) k_select(PaCfg cfg, PaBatchDev b, PaSelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    unsigned long long* s_sort = (unsigned long long*)smem_raw + (size_t)wib * PA_SORTCAP;
    const int64_t gw = (int64_t)blockIdx.x * wpb + wib, nw = (int64_t)gridDim.x * wpb;
    const float INF = __int_as_float(0x7f800000);

    for (int64_t p = gw; p < a.n_psm; p += nw) {
        const int status = a.psm_status[p];
        const int k = b.n_mod[p];
        const int64_t mo = a.mod_off[p];
        const int S = a.psm_S[p];
        const int64_t I = a.iso_off[p + 1] - a.iso_off[p];
        const int64_t ib = a.iso_off[p];
        if (lane == 0) {
            if (a.psm_status_out) a.psm_status_out[p] = status;
            if (a.n_iso) a.n_iso[p] = I;
            if (a.n_sites) a.n_sites[p] = S;
        }
        if (status != PA_PSM_OK || I == 0 || k >= S) {
            // no isoform: best_sequence "" / best_score -1 (cpp/Ascore.cpp:273-295); k >= #sites is
            // "unambiguous" (cpp/Ascore.cpp:38-51: ascores inf, no alternatives); errors report NaN
            const bool ok = status == PA_PSM_OK;
            const float fill = ok ? INF : __int_as_float(0x7fc00000);
            if (lane == 0) {
                uint64_t sig = 0; float sc = ok ? -1.f : fill;
                if (ok && I > 0) { sig = (S >= 64) ? ~0ull : ((1ull << S) - 1ull); sc = a.iso.w[ib]; }
                if (a.best_sig) a.best_sig[p] = sig;
                if (a.best_score) a.best_score[p] = sc;
                a.best_idx[p] = (ok && I > 0) ? 0u : 0xffffffffu;
            }
            for (int j = lane; j < k; j += 32) {
                if (a.ascores) a.ascores[mo + j] = fill;
                if (a.alt_sites) a.alt_sites[mo + j] = 0;
                a.mod_psm[mo + j - a.mod_lo] = -1;
            }
            continue;
        }
        // ---- best isoform ---------------------------------------------------------------
        float wmax = -INF;
        for (int64_t q = lane; q < I; q += 32) { float w = a.iso.w[ib + q]; wmax = w > wmax ? w : wmax; }
        for (int o = 16; o > 0; o >>= 1) { float t = __shfl_xor_sync(PA_FULL, wmax, o); wmax = t > wmax ? t : wmax; }
        int ties = 0;
        uint32_t first_tie = 0xffffffffu;
        for (int64_t q = lane; q < I; q += 32)
            if (a.iso.w[ib + q] == wmax) { ties++; if (first_tie == 0xffffffffu) first_tie = (uint32_t)q; }
        for (int o = 16; o > 0; o >>= 1) {
            ties += __shfl_xor_sync(PA_FULL, ties, o);
            uint32_t t = __shfl_xor_sync(PA_FULL, first_tie, o); first_tie = t < first_tie ? t : first_tie;
        }
        uint32_t best = first_tie;
        if (ties > 1) {
            const int64_t po = a.perm_off[S * 64 + k];
            const uint32_t* perm = a.perm_pool + po;
            if (I <= 16) {
                // std::sort on <= 16 elements is a stable insertion sort: the first maximal element
                // in hash-iteration order stays in front
                uint32_t id = lane < I ? perm[lane] : 0u;
                bool is = lane < I && a.iso.w[ib + id] == wmax;
                unsigned bal = __ballot_sync(PA_FULL, is);
                best = __shfl_sync(PA_FULL, id, __ffs(bal) - 1);
            } else {
                unsigned long long* arr = (I <= PA_SORTCAP) ? s_sort : a.g_sort + ib;
                __syncwarp();
                for (int64_t q = lane; q < I; q += 32) {
                    uint32_t id = perm[q];
                    arr[q] = ((unsigned long long)(uint32_t)__float_as_int(a.iso.w[ib + id]) << 32) | id;
                }
                __syncwarp();
                if (lane == 0) best = pa_gcc_sort_front(arr, (long)I);
                best = __shfl_sync(PA_FULL, best, 0);
                __syncwarp();
            }
        }
        const uint64_t best_bits = pa_unrank(cfg.binom, S, k, best);
        if (lane == 0) {
            if (a.best_sig) a.best_sig[p] = best_bits;
            if (a.best_score) a.best_score[p] = a.iso.w[ib + best];
            a.best_idx[p] = best;
        }
        // ---- tied best competitors per modified site (cpp/Ascore.cpp:212-254) ---------------
        const uint64_t all = (S >= 64) ? ~0ull : ((1ull << S) - 1ull);
        const uint64_t free_sites = all & ~best_bits;
        uint64_t rem = best_bits;
        for (int j = 0; j < k; j++) {
            const int site = __ffsll((long long)rem) - 1;
            rem &= rem - 1;
            float w0 = -INF, w1 = -INF;      // competitor scores of free sites lane, lane+32
            if (lane < S && ((free_sites >> lane) & 1ull))
                w0 = a.iso.w[ib + pa_rank(cfg.binom, S, k, (best_bits & ~(1ull << site)) | (1ull << lane))];
            if (lane + 32 < S && ((free_sites >> (lane + 32)) & 1ull))
                w1 = a.iso.w[ib + pa_rank(cfg.binom, S, k, (best_bits & ~(1ull << site)) | (1ull << (lane + 32)))];
            float m = w0 > w1 ? w0 : w1;
            for (int o = 16; o > 0; o >>= 1) { float t = __shfl_xor_sync(PA_FULL, m, o); m = t > m ? t : m; }
            const bool f0 = lane < S && ((free_sites >> lane) & 1ull) && w0 == m;
            const bool f1 = lane + 32 < S && ((free_sites >> (lane + 32)) & 1ull) && w1 == m;
            const uint64_t tie = (uint64_t)__ballot_sync(PA_FULL, f0) | ((uint64_t)__ballot_sync(PA_FULL, f1) << 32);
            if (lane == 0) {
                if (a.alt_sites) a.alt_sites[mo + j] = tie;
                a.tie[mo + j - a.mod_lo] = tie;
                a.mod_psm[mo + j - a.mod_lo] = (int32_t)p;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// K3b: one THREAD per (PSM, modified site) entry.  For every tied competitor: the site-determining
// ion comparison of cpp/Ascore.cpp:157-210 + cpp/ModifiedPeptide.cpp:259-320, computed without
// materialising or sorting the fragment lists: each (charge, neutral-loss variant) is a stream that
// is already ascending along the peptide, so "sort, then tolerance-merge" becomes a k-way streaming
// merge whose element VALUES are exactly those of the reference's sorted lists.  If a stream is
// ever found non-ascending (non-positive residue mass), or the shape is out of range (too many
// streams, one-residue peptide), the entry is queued for k_ascore_generic instead.
// ---------------------------------------------------------------------------------------------
#define PA_MAXSTREAM 12

struct PaAscArgs {
    int64_t n_entries;           // mod entries of this chunk
    int64_t mod_lo;              // absolute index of the chunk's first entry
    const int32_t* mod_psm;      // [n_entries] chunk-relative PSM, -1 = nothing to do
    const unsigned long long* tie;   // [n_entries] tied competitor sites
    const uint32_t* best_idx;    // [n_psm]
    const int64_t* mod_off;      // [n_psm+1] absolute
    const int64_t* iso_off;
    const int32_t* psm_S;
    PaIso iso;
    float* ascores;              // absolute-indexed output (may be null)
    int32_t* generic_list;       // entries that need the generic kernel
    int* generic_count;
};

struct AscPep {                  // what a thread needs to know about its peptide
    const uint8_t* pep;
    int L, Z, a0, a1;
    const uint32_t* aux_pos;
    const float* aux_mass;
};

// residue mass / neutral-loss index of residue i in modification state st
// (cpp/ModifiedPeptide.cpp:24-79 evaluated on the fly instead of tabulated)
__device__ __forceinline__ float asc_res(const PaCfg& cfg, const AscPep& q, int i, int st, int& nlidx) {
    const int c = (int)q.pep[i] - 'A';
    float m = __ldg(cfg.res_tab + c);
    if (st) m = __fadd_rn(m, cfg.mod_mass);
    nlidx = 0;
    if (cfg.has_nl) nlidx = st ? __ldg(cfg.nl_lo_tab + c) : __ldg(cfg.nl_up_tab + c);
    for (int a = q.a0; a < q.a1; a++) {
        const uint32_t pos = q.aux_pos[a];
        const int idx = pos > 0 ? (int)pos - 1 : 0;
        if (idx == i) {
            m = __fadd_rn(m, q.aux_mass[a]);
            if (cfg.has_nl && __ldg(cfg.nl_lo_tab + c)) nlidx = __ldg(cfg.nl_lo_tab + c);
        }
    }
    return m;
}

struct AscList {                 // streams of one isoform for one ion type
    float run[PA_MAXSTREAM];
    float val[PA_MAXSTREAM];
    float sigma[PA_MAXSTREAM];
    short step[PA_MAXSTREAM];
    short zq[PA_MAXSTREAM];
    int nq;                      // number of streams
    int left;                    // elements not yet popped
};

// returns false when the shape is not supported by the streaming formulation
__device__ __forceinline__ bool asc_init(const PaCfg& cfg, const AscPep& q, uint64_t mlo, uint64_t mhi, bool fwd,
                                         double a1, double a2, AscList& ls) {
    const int L = q.L, Z = q.Z, steps = L - 1;
    float sig[PA_MAXSTREAM], run_at[PA_MAXSTREAM];
    short start[PA_MAXSTREAM];
    int V = 1;
    if (!cfg.has_nl) {
        // one stream per charge: no loss (sigma 0), present from the first residue on
        if (Z > PA_MAXSTREAM) return false;
        const int i = fwd ? 0 : L - 1;
        const int st = (int)(((i < 64) ? (mlo >> i) : (mhi >> (i - 64))) & 1ull);
        int idx;
        sig[0] = 0.f; start[0] = 0;
        run_at[0] = asc_res(cfg, q, i, st, idx);
    } else {
        // pass 1: the final neutral-loss state says which sums will ever exist
        int nls = 0;
        for (int step = 0; step < steps; step++) {
            const int i = fwd ? step : L - 1 - step;
            const int st = (int)(((i < 64) ? (mlo >> i) : (mhi >> (i - 64))) & 1ull);
            int idx;
            (void)asc_res(cfg, q, i, st, idx);
            if (idx) nls = pa_nl_bump(nls, idx);
        }
        V = cfg.nl_nvar[nls];
        if (V * Z > PA_MAXSTREAM) return false;
        for (int v = 0; v < V; v++) { sig[v] = __ldg(cfg.nl_sums + nls * 16 + v); start[v] = -1; run_at[v] = 0.f; }
        // pass 2: the step at which each sum first becomes available (the stack only grows, so
        // it stays available afterwards) and the running sum there
        int started = 0;
        float run = 0.f;
        nls = 0;
        for (int step = 0; step < steps && started < V; step++) {
            const int i = fwd ? step : L - 1 - step;
            const int st = (int)(((i < 64) ? (mlo >> i) : (mhi >> (i - 64))) & 1ull);
            int idx;
            const float r = asc_res(cfg, q, i, st, idx);
            run = (step == 0) ? r : __fadd_rn(r, run);
            const int before = nls;
            if (idx) nls = pa_nl_bump(nls, idx);
            if (step == 0 || nls != before) {
                const int nv = cfg.nl_nvar[nls];
                for (int v = 0; v < V; v++) {
                    if (start[v] >= 0) continue;
                    bool in = false;
                    for (int u = 0; u < nv; u++) in |= (__ldg(cfg.nl_sums + nls * 16 + u) == sig[v]);
                    if (in) { start[v] = (short)step; run_at[v] = run; started++; }
                }
            }
        }
    }
    ls.nq = 0; ls.left = 0;
    for (int v = 0; v < V; v++) {
        if (cfg.has_nl && start[v] < 0) continue;     // cannot happen: every final sum appears somewhere
        for (int z = 1; z <= Z; z++) {
            const int qi = ls.nq++;
            ls.run[qi] = run_at[v]; ls.sigma[qi] = sig[v]; ls.step[qi] = start[v]; ls.zq[qi] = (short)z;
            const double d = __dsub_rn(__dadd_rn((double)__fsub_rn(run_at[v], sig[v]), a1), a2);
            ls.val[qi] = pa_charge_mz(d, z);
            ls.left += steps - start[v];
        }
    }
    return true;
}

// pop the smallest pending fragment of the list; `mono` is cleared if a stream ever decreases
__device__ __forceinline__ float asc_pop(const PaCfg& cfg, const AscPep& q, uint64_t mlo, uint64_t mhi, bool fwd,
                                         double a1, double a2, AscList& ls, bool& mono) {
    int bq = 0;
    float x = ls.val[0];
    for (int i = 1; i < ls.nq; i++) { const float v = ls.val[i]; if (v < x) { x = v; bq = i; } }
    const int L = q.L, steps = L - 1;
    const int step = ls.step[bq] + 1;
    ls.step[bq] = (short)step;
    ls.left--;
    if (step < steps) {
        const int i = fwd ? step : L - 1 - step;
        const int st = (int)(((i < 64) ? (mlo >> i) : (mhi >> (i - 64))) & 1ull);
        int idx;
        const float r = asc_res(cfg, q, i, st, idx);
        const float run = __fadd_rn(r, ls.run[bq]);
        ls.run[bq] = run;
        const double d = __dsub_rn(__dadd_rn((double)__fsub_rn(run, ls.sigma[bq]), a1), a2);
        const float nv = pa_charge_mz(d, ls.zq[bq]);
        if (nv < x) mono = false;
        ls.val[bq] = nv;
    } else ls.val[bq] = __int_as_float(0x7f800000);
    return x;
}

__device__ __forceinline__ int asc_match(const float* pm, const uint8_t* pr, int R, const uint8_t* ctab, float cbase,
                                         float cinv, float f, float err, int err_gt_half) {
    const float lo = __fsub_rn(f, err), hi = __fadd_rn(f, err);
    int a;
    if (cinv != 0.f) {
        a = __ldg(ctab + pa_cell(lo, cbase, cinv));
        while (a < R && !(__ldg(pm + a) > lo)) a++;
    } else {
        a = 0;
        int n = R;
        while (n > 0) {
            int h = n >> 1;
            if (!(__ldg(pm + a + h) > lo)) { a += h + 1; n -= h + 1; } else n = h;
        }
    }
    int best = 255;
    for (; a < R; a++) {
        const float p = __ldg(pm + a);
        if (!(p < hi)) break;
        if (err_gt_half && !((double)f >= (double)p - .5)) continue;
        const int r = __ldg(pr + a);
        best = r < best ? r : best;
    }
    return best;
}

__global__ void __launch_bounds__(128) k_ascore(PaCfg cfg, PaBatchDev b, PaAscArgs a) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_entries) return;
    const int32_t p = a.mod_psm[t];
    if (p < 0) return;
    const int j = (int)(a.mod_lo + t - a.mod_off[p]);
    const int S = a.psm_S[p], k = b.n_mod[p];
    const int64_t ib = a.iso_off[p];
    AscPep q;
    const int po = b.pep_off[p];
    q.pep = b.pep + po; q.L = b.pep_off[p + 1] - po; q.Z = b.max_charge[p];
    q.a0 = 0; q.a1 = 0; q.aux_pos = b.aux_pos; q.aux_mass = b.aux_mass;
    if (b.aux_off != nullptr) { q.a0 = b.aux_off[p]; q.a1 = b.aux_off[p + 1]; }
    const int sp = b.psm_spec[p];
    const int64_t off = b.spec_off[sp] - b.spec_base;
    const float* pm = b.rmz + off;
    const uint8_t* pr = b.rrank + off;
    const int R = b.rcount[sp];
    const uint8_t* ctab = b.ctab + (size_t)sp * PA_NCELL;
    const float2 chead = b.chead[sp];

    const uint32_t best = a.best_idx[p];
    const uint64_t best_bits = pa_unrank(cfg.binom, S, k, best);
    uint64_t rem = best_bits;
    for (int jj = 0; jj < j; jj++) rem &= rem - 1;
    const int site = __ffsll((long long)rem) - 1;
    // site index -> residue position (sites are the modifiable residues N->C)
    int site_res[64];
    {
        int n = 0;
        for (int i = 0; i < q.L && n < 64; i++) {
            const int c = (int)q.pep[i] - 'A';
            const bool is = ((cfg.mod_letters >> c) & 1u) || (cfg.allow_n && i == 0) || (cfg.allow_c && i == q.L - 1);
            if (is) site_res[n++] = i;
        }
    }
    uint64_t alo = 0, ahi = 0;
    for (uint64_t bb = best_bits; bb; bb &= bb - 1) {
        const int pos = site_res[__ffsll((long long)bb) - 1];
        if (pos < 64) alo |= 1ull << pos; else ahi |= 1ull << (pos - 64);
    }
    float scA[PA_N_TOP];
    pa_depth_scores(cfg, a.iso.lo[ib + best], a.iso.hi[ib + best], (int)a.iso.nfrag[ib + best], scA);
    const float wA = a.iso.w[ib + best];

    bool generic = (q.L < 2);
    float asc = __int_as_float(0x7f800000);
    for (uint64_t tt = a.tie[t]; tt && !generic; tt &= tt - 1) {
        const int u = __ffsll((long long)tt) - 1;
        const uint64_t cb = (best_bits & ~(1ull << site)) | (1ull << u);
        const uint32_t ci = pa_rank(cfg.binom, S, k, cb);
        const float wB = a.iso.w[ib + ci];
        float amb = 0.f;
        if (!((double)fabsf(__fsub_rn(wA, wB)) < 1e-6)) {
            float scB[PA_N_TOP];
            pa_depth_scores(cfg, a.iso.lo[ib + ci], a.iso.hi[ib + ci], (int)a.iso.nfrag[ib + ci], scB);
            float max_diff = 0.f;
            int depth = 0;
#pragma unroll
            for (int d = 0; d < PA_N_TOP; d++) {
                const float diff = __fsub_rn(scA[d], scB[d]);
                if (diff > max_diff) { max_diff = diff; depth = d; }
            }
            // competitor mask = best mask with the mod moved from `site` to site u
            uint64_t blo = alo, bhi = ahi;
            { const int pos = site_res[site]; if (pos < 64) blo &= ~(1ull << pos); else bhi &= ~(1ull << (pos - 64)); }
            { const int pos = site_res[u]; if (pos < 64) blo |= 1ull << pos; else bhi |= 1ull << (pos - 64); }
            int hitsA = 0, hitsB = 0, trialsA = 0, trialsB = 0;
            for (int ti = 0; ti < cfg.n_types && !generic; ti++) {
                const char type = cfg.types[ti];
                const bool fwd = (type == 'b' || type == 'c');
                double a1, a2;
                pa_type_consts(type, a1, a2);
                AscList A, B;
                if (!asc_init(cfg, q, alo, ahi, fwd, a1, a2, A) || !asc_init(cfg, q, blo, bhi, fwd, a1, a2, B)) { generic = true; break; }
                bool mono = true;
                float x = 0.f, y = 0.f;
                bool hx = false, hy = false;         // a popped element is pending
                // greedy tolerance merge of cpp/ModifiedPeptide.cpp:288-316 over the two streams
                for (;;) {
                    if (!hx && A.left > 0) { x = asc_pop(cfg, q, alo, ahi, fwd, a1, a2, A, mono); hx = true; }
                    if (!hy && B.left > 0) { y = asc_pop(cfg, q, blo, bhi, fwd, a1, a2, B, mono); hy = true; }
                    if (!hx && !hy) break;
                    if (!hy) { trialsA++; hitsA += asc_match(pm, pr, R, ctab, chead.x, chead.y, x, cfg.err, cfg.err_gt_half) <= depth; hx = false; }
                    else if (!hx) { trialsB++; hitsB += asc_match(pm, pr, R, ctab, chead.x, chead.y, y, cfg.err, cfg.err_gt_half) <= depth; hy = false; }
                    else if (fabsf(__fsub_rn(x, y)) < cfg.err) { hx = false; hy = false; }
                    else if (x < y) { trialsA++; hitsA += asc_match(pm, pr, R, ctab, chead.x, chead.y, x, cfg.err, cfg.err_gt_half) <= depth; hx = false; }
                    else { trialsB++; hitsB += asc_match(pm, pr, R, ctab, chead.x, chead.y, y, cfg.err, cfg.err_gt_half) <= depth; hy = false; }
                }
                if (!mono) generic = true;
            }
            if (!generic) {
                const float sA = __ldg(cfg.T + pa_tab_index(trialsA, hitsA, depth));
                const float sB = __ldg(cfg.T + pa_tab_index(trialsB, hitsB, depth));
                amb = __fsub_rn(sA, sB);
            }
        }
        asc = amb < asc ? amb : asc;
    }
    if (generic) {
        const int slot = atomicAdd(a.generic_count, 1);
        a.generic_list[slot] = (int32_t)t;
        return;
    }
    if (a.ascores) a.ascores[a.mod_lo + t] = asc;
}

// K3c: generic (warp-cooperative, list-materialising) Ascore for the entries k_ascore queued.
__global__ void __launch_bounds__(256) k_ascore_generic(PaCfg cfg, PaBatchDev b, PaAscArgs a, float* g_lists,
                                                         int64_t list_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    SelSmem* sm = (SelSmem*)smem_raw + wib;
    const int64_t gw = (int64_t)blockIdx.x * wpb + wib, nw = (int64_t)gridDim.x * wpb;
    const int n = *a.generic_count;
    for (int64_t e = gw; e < n; e += nw) {
        const int64_t t = a.generic_list[e];
        const int32_t p = a.mod_psm[t];
        const int j = (int)(a.mod_lo + t - a.mod_off[p]);
        const int S = a.psm_S[p], k = b.n_mod[p];
        const int64_t ib = a.iso_off[p];
        PsmInfo info;
        pa_setup_psm(cfg, b, p, &sm->psm, info, true);
        const uint32_t best = a.best_idx[p];
        const uint64_t best_bits = pa_unrank(cfg.binom, S, k, best);
        uint64_t rem = best_bits;
        for (int jj = 0; jj < j; jj++) rem &= rem - 1;
        const int site = __ffsll((long long)rem) - 1;
        float scBest[PA_N_TOP];
        pa_depth_scores(cfg, a.iso.lo[ib + best], a.iso.hi[ib + best], (int)a.iso.nfrag[ib + best], scBest);
        const float wbest = a.iso.w[ib + best];
        float* raw0 = sm->raw[0]; float* raw1 = sm->raw[1]; float* srt0 = sm->srt[0]; float* srt1 = sm->srt[1];
        long long per_type = (long long)(info.L > 1 ? info.L - 1 : 1) * cfg.nvar_cap * info.Z;
        if (per_type > PA_LCAP) {     // list_stride = power of two >= the chunk's longest list
            float* g = g_lists + (size_t)gw * 4 * list_stride;
            raw0 = g; raw1 = g + list_stride; srt0 = g + 2 * list_stride; srt1 = g + 3 * list_stride;
        }
        float asc = __int_as_float(0x7f800000);
        for (uint64_t tt = a.tie[t]; tt; tt &= tt - 1) {
            const int u = __ffsll((long long)tt) - 1;
            const uint64_t cb = (best_bits & ~(1ull << site)) | (1ull << u);
            const uint32_t ci = pa_rank(cfg.binom, S, k, cb);
            float scC[PA_N_TOP];
            pa_depth_scores(cfg, a.iso.lo[ib + ci], a.iso.hi[ib + ci], (int)a.iso.nfrag[ib + ci], scC);
            const float amb = pa_ambiguity(cfg, sm, info, best_bits, scBest, wbest, cb, scC, a.iso.w[ib + ci],
                                           raw0, raw1, srt0, srt1);
            asc = amb < asc ? amb : asc;
        }
        if (lane == 0 && a.ascores) a.ascores[a.mod_lo + t] = asc;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// export of one PSM's per-isoform table (pep_scores) and a stand-alone ambiguity call
// ---------------------------------------------------------------------------------------------
__global__ void k_export_psm(PaCfg cfg, int64_t ib, int64_t I, int S, int k, PaIso iso, uint64_t* sig,
                             int32_t* counts, float* scores, float* weighted, int32_t* total) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= I) return;
    unsigned long long lo = iso.lo[ib + q], hi = iso.hi[ib + q];
    int n = (int)iso.nfrag[ib + q];
    sig[q] = pa_unrank(cfg.binom, S, k, (uint32_t)q);
    for (int d = 0; d < PA_N_TOP; d++) {
        int c = pa_cum_get(lo, hi, d);
        counts[q * PA_N_TOP + d] = c;
        scores[q * PA_N_TOP + d] = cfg.T[pa_tab_index(n, c, d)];
    }
    weighted[q] = iso.w[ib + q];
    total[q] = n;
}

struct PaAmbArgs {
    int64_t psm;
    uint64_t sigA, sigB;
    float scA[PA_N_TOP], scB[PA_N_TOP];
    float wA, wB;
    float* out;
    float* g_lists;
    int64_t list_stride;
};

__global__ void __launch_bounds__(32) k_ambiguity(PaCfg cfg, PaBatchDev b, PaAmbArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SelSmem* sm = (SelSmem*)smem_raw;
    PsmInfo info;
    pa_setup_psm(cfg, b, a.psm, &sm->psm, info, true);
    float* raw0 = sm->raw[0]; float* raw1 = sm->raw[1]; float* srt0 = sm->srt[0]; float* srt1 = sm->srt[1];
    long long per_type = (long long)(info.L > 1 ? info.L - 1 : 1) * cfg.nvar_cap * info.Z;
    if (per_type > PA_LCAP) {
        float* g = a.g_lists;
        raw0 = g; raw1 = g + a.list_stride; srt0 = g + 2 * a.list_stride; srt1 = g + 3 * a.list_stride;
    }
    float r = pa_ambiguity(cfg, sm, info, a.sigA, a.scA, a.wA, a.sigB, a.scB, a.wB, raw0, raw1, srt0, srt1);
    if ((threadIdx.x & 31) == 0) *a.out = r;
}
