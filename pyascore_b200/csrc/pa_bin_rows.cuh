// K1 "row form": BinnedSpectra for the spectra that are the rule -- m/z ascending, at most PA_ROWS_MAXCAP peaks, a sane
// m/z range, n_top <= 31.  cpp/Spectra.cpp:43-68 (bounds + bin index), :24-41 (top n_top of every bin by intensity).
//
// One warp per spectrum, three passes of one peak per lane and 32 peaks per round, nothing staged: a round of the binning
// pass reads its 32 m/z and intensities straight from global memory (coalesced 8- or 4-byte loads, the next round's already
// in flight), and leaves (float)m/z, the ranking key (float)intensity and the bin in shared memory; the ranking pass is the
// all-pairs count inside the bin's contiguous run [start, end) (four keys per LDS.128); the output pass compacts the kept
// peaks with one ballot per round.  Shared memory is addressed with 32-bit shared-window addresses through ld/st.shared
// (the generic-pointer form recomputes the window base at every access), full rounds carry no bounds predicates (the last,
// partial round is a second instantiation of the same body), and "key > mine" costs 1.5 instructions (pa_gt_bits).
//
// Everything unusual is DECLINED, not handled: the spectrum's index goes to a list and k_bin_topn (its exact general
// paths included) runs over that list afterwards.  Declined: more peaks than the slot holds, m/z not ascending as float64
// (or NaN), ends outside [0, 1e6], more than PA_NBIN_SMEM bins, and -- in the instantiation for host-narrowed m/z -- the
// spectra that kept an exact float64 copy.  Ties on the float keys are handled here, as in k_bin_topn: the bin's mask of
// taken ranks exposes them and the bin is re-ranked on the exact keys from global memory.
#pragma once
#include <type_traits>

#define PA_ROWS_MAXCAP 512
// per warp: key f32[cap + 4] | mzf f32[cap] | bin u8[cap] | cnt u8[cap] | range u32[132] | tie mask u32[128] | cell u8[256]
#define PA_ROWS_RANGE_BYTES ((PA_NBIN_SMEM + 4) * 4)
#define PA_ROWS_SLOT_BYTES(cap) ((size_t)(cap) * 10 + 16 + PA_ROWS_RANGE_BYTES + PA_NBIN_SMEM * 4 + PA_NCELL)

// Counting "key > mine" at 1.5 instructions per key: FSET.BF leaves the BITS of 1.0f (0x3f800000 = 127 << 23) or 0, and the
// integer sum of n such words is n * 127 << 23 modulo 2^32, from which n < 512 comes back as ((sum >> 23) * 383) & 511
// (383 is the inverse of 127 modulo 512).  Four keys cost four FSET.BF and two three-input integer adds.
__device__ __forceinline__ int pa_gt_bits(float a, float b) {
    float d;
    asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return __float_as_int(d);
}
#define PA_GT4(v, hi) (pa_gt_bits((v).x, hi) + pa_gt_bits((v).y, hi) + pa_gt_bits((v).z, hi) + pa_gt_bits((v).w, hi))
// the same for the first / last group of a run: only the keys at positions r .. r + 3 that fall inside [0, n) count
#define PA_GT4_IN(v, hi, r, n)                                                                          \
    (((unsigned)(r) < (n) ? pa_gt_bits((v).x, hi) : 0) + ((unsigned)((r) + 1) < (n) ? pa_gt_bits((v).y, hi) : 0) + \
     ((unsigned)((r) + 2) < (n) ? pa_gt_bits((v).z, hi) : 0) + ((unsigned)((r) + 3) < (n) ? pa_gt_bits((v).w, hi) : 0))
__device__ __forceinline__ int pa_gt_count(int sum) { return (int)((((unsigned)sum >> 23) * 383u) & 511u); }

// shared memory through 32-bit shared-window addresses
__device__ __forceinline__ void pa_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void pa_stsf(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void pa_sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v)); }
__device__ __forceinline__ void pa_sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void pa_sts64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v)); }
__device__ __forceinline__ void pa_sts128z(uint32_t a) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0)); }
__device__ __forceinline__ uint32_t pa_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float pa_ldsf(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t pa_lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned long long pa_lds64(uint32_t a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 pa_lds128f(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 pa_lds128u(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t pa_atoms_or(uint32_t a, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}

// F32: float32 intensities (pa_batch.inten32).  NARROW: float32 m/z from the host's narrowing pass (pa_narrow_mz).
template <bool F32, bool NARROW>
__global__ void __launch_bounds__(256, 4) k_bin_rows(PaBinArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename std::conditional<NARROW, float, double>::type mz_t;
    typedef typename std::conditional<F32, float, double>::type in_t;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int cap = a.cap;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw) + (uint32_t)wib * (uint32_t)PA_ROWS_SLOT_BYTES(cap);
    const uint32_t S_KEY = sb;                                  // ranking keys; after the ranking pass the kept m/z
    const uint32_t S_MZF = sb + (uint32_t)cap * 4 + 16;
    const uint32_t S_BIN = sb + (uint32_t)cap * 8 + 16;
    const uint32_t S_CNT = S_BIN + (uint32_t)cap;
    const uint32_t S_RANGE = S_CNT + (uint32_t)cap;             // start | end << 16 of every bin's run; entry 128 is a dummy
    const uint32_t S_BMASK = S_RANGE + PA_ROWS_RANGE_BYTES;     // per bin: ranks taken (bit r), bit 31 = tie seen
    const uint32_t S_CELL = S_BMASK + PA_NBIN_SMEM * 4;
    const int64_t gw = (int64_t)blockIdx.x * wpb + wib, nw = (int64_t)gridDim.x * wpb;
    const int n_top = a.n_top;
    const unsigned below = (1u << lane) - 1u;
    const double dbs = (double)a.bin_size;
    const double inv_bs = __ddiv_rn(1.0, dbs);

    int64_t o0 = 0, o1 = 0;
    if (gw < a.n_spec) { o0 = a.spec_off[gw]; o1 = a.spec_off[gw + 1]; }
    for (int64_t s = gw; s < a.n_spec; s += nw) {
        const int64_t off = o0 - a.peak_base, Pl = o1 - o0;
        if (s + nw < a.n_spec) { o0 = a.spec_off[s + nw]; o1 = a.spec_off[s + nw + 1]; }     // the next spectrum's, early
        if (Pl <= 0) { if (lane == 0) { a.rcount[s] = 0; a.chead[s] = make_float2(0.f, 0.f); } continue; }
        auto decline = [&]() { if (lane == 0) a.list[atomicAdd(a.list_n, 1u)] = (int32_t)s; };
        if (Pl > cap) { decline(); continue; }
        if (NARROW) { if (a.esc_off[s] >= 0) { decline(); continue; } }
        const int P = (int)Pl;
        const mz_t* __restrict__ mzp = (NARROW ? (const mz_t*)a.mz32 : (const mz_t*)a.mz) + off;
        const in_t* __restrict__ inp = (F32 ? (const in_t*)a.inten32 : (const in_t*)a.inten) + off;
        // one round's loads: m/z of the peak and of its predecessor (the order test), intensity
        mz_t mA = 0, pA = 0, mB = 0, pB = 0;
        in_t kA = 0, kB = 0;
        if (lane < P) { mA = mzp[lane]; pA = mzp[lane > 0 ? lane - 1 : 0]; kA = inp[lane]; }
        const double mn = (double)mzp[0], mx = (double)mzp[P - 1];
        if (!(mn >= 0. && mx <= 1e6)) { decline(); continue; }          // (NaN ends fail both)
        // cpp/Spectra.cpp:46-48: the 100 is a literal there, independent of bin_size
        const float min_mz = __double2float_rn(__dmul_rn(floor(__ddiv_rn(mn, 100.)), 100.));
        const float max_mz = __double2float_rn(__dmul_rn(ceil(__ddiv_rn(mx, 100.)), 100.));
        int n_bins = (int)ceilf(__fdiv_rn(__fsub_rn(max_mz, min_mz), a.bin_size));
        if (n_bins < 1) n_bins = 1;
        if (n_bins > PA_NBIN_SMEM) { decline(); continue; }
        const int top_bin = n_bins - 1;
        const double dmin = (double)min_mz;
        pa_sts128z(S_BMASK + 16 * lane);

        // ---- binning ----
        bool sorted = true;
        int carry_bin = PA_NBIN_SMEM;                 // "no bin yet": its end lands in the dummy entry
        int bq = 0;
        auto row = [&](auto tail, int i, mz_t m, mz_t mp, in_t kin) {
            constexpr bool TAIL = decltype(tail)::value;
            const bool valid = !TAIL || i < P;
            // floor((m - min) / bin_size) as the reference computes it; the reciprocal product decides unless it
            // lands within 1e-9 of an integer, where the IEEE quotient is taken
            const double x = __dsub_rn((double)m, dmin);
            const double t = __dmul_rn(x, inv_bs);
            double q = floor(t);
            const double fr = __dsub_rn(t, q);
            if (!(fr > 1e-9 && fr < 1. - 1e-9)) q = floor(__ddiv_rn(x, dbs));
            bq = min(max((int)q, 0), top_bin);                           // (the conversion saturates; NaN -> 0)
            if (TAIL && !valid) bq = PA_NBIN_SMEM;
            if (valid) {
                sorted = sorted && (m >= mp);                            // NaN fails
                pa_stsf(S_MZF + 4 * i, NARROW ? (float)m : __double2float_rn((double)m));
                pa_stsf(S_KEY + 4 * i, F32 ? (float)kin : __double2float_rn((double)kin));
                pa_sts8(S_BIN + i, (uint32_t)bq);
            }
            int bprev = __shfl_up_sync(PA_FULL, bq, 1);
            if (lane == 0) bprev = carry_bin;
            if (valid && bq != bprev) {
                pa_sts16(S_RANGE + 4 * bq, (uint32_t)i);
                pa_sts16(S_RANGE + 4 * bprev + 2, (uint32_t)i);
            }
            carry_bin = __shfl_sync(PA_FULL, bq, 31);
        };
        for (int base = 0;;) {
            // A holds round `base`; the loads of the next round go out before A is worked on
            int i = base + lane, in = i + 32;
            if (in < P) { mB = mzp[in]; pB = mzp[in - 1]; kB = inp[in]; }
            if (base + 32 <= P) row(std::false_type(), i, mA, pA, kA); else row(std::true_type(), i, mA, pA, kA);
            base += 32;
            if (base >= P) break;
            i = base + lane; in = i + 32;
            if (in < P) { mA = mzp[in]; pA = mzp[in - 1]; kA = inp[in]; }
            if (base + 32 <= P) row(std::false_type(), i, mB, pB, kB); else row(std::true_type(), i, mB, pB, kB);
            base += 32;
            if (base >= P) break;
        }
        if (lane == ((P - 1) & 31)) pa_sts16(S_RANGE + 4 * bq + 2, (uint32_t)P);       // the last run ends with the spectrum
        if (!__all_sync(PA_FULL, sorted)) { decline(); __syncwarp(); continue; }
        __syncwarp();

        // ---- ranking: peaks of the same bin that beat this one, counted on the float keys ----
        // Two peaks of a bin with the same key get the same count: every kept peak marks its count in the bin's mask,
        // and a mark found already set flags the bin for exact ranking.
        for (int base = 0; base < P; base += 32) {
            const int i = base + lane;
            if (i < P) {
                const uint32_t bqi = pa_lds8(S_BIN + i);
                const float hi = pa_ldsf(S_KEY + 4 * i);
                const uint32_t rg = pa_lds32(S_RANGE + 4 * bqi);
                const int b0 = (int)(rg & 0xffffu), b1 = (int)(rg >> 16);
                const unsigned n = (unsigned)(b1 - b0);
                const int a0 = b0 & ~3, al = (b1 - 1) & ~3;              // first peak of the first / last group of four
                uint32_t ga = S_KEY + 4 * a0;
                const uint32_t gend = S_KEY + 4 * al;
                float4 v = pa_lds128f(ga);
                int acc = PA_GT4_IN(v, hi, a0 - b0, n);
#pragma unroll 1
                for (ga += 16; ga < gend; ga += 16) {
                    v = pa_lds128f(ga);
                    acc += PA_GT4(v, hi);
                }
                if (al > a0) {
                    v = pa_lds128f(gend);
                    acc += PA_GT4_IN(v, hi, al - b0, n);
                }
                int c = pa_gt_count(acc);
                if (c < n_top) {
                    const uint32_t bit = 1u << c;
                    if (pa_atoms_or(S_BMASK + 4 * bqi, bit) & bit) pa_atoms_or(S_BMASK + 4 * bqi, 0x80000000u);
                } else c = 255;
                pa_sts8(S_CNT + i, (uint32_t)c);
            }
        }
        __syncwarp();
        const uint4 tm = pa_lds128u(S_BMASK + 16 * lane);
        const bool ties = __any_sync(PA_FULL, ((tm.x | tm.y | tm.z | tm.w) >> 31) != 0u);

        // ---- output: kept peaks in m/z order as {mz, rank}; their m/z also over the keys, for the cell index ----
        float2* __restrict__ rpk = a.rpk + off;
        int out = 0;
        for (int base = 0; base < P; base += 32) {
            const int i = base + lane;
            int cnt = 255;
            float mzf = 0.f;
            if (i < P) {
                cnt = (int)pa_lds8(S_CNT + i);
                mzf = pa_ldsf(S_MZF + 4 * i);
                if (ties) {
                    const uint32_t bqi = pa_lds8(S_BIN + i);
                    if (pa_lds32(S_BMASK + 4 * bqi) >> 31) {
                        // exact rank from the full intensity keys (read back from global memory); an equal intensity
                        // wins only from an earlier index
                        const uint32_t rg = pa_lds32(S_RANGE + 4 * bqi);
                        const int b0 = (int)(rg & 0xffffu), b1 = (int)(rg >> 16);
                        const uint64_t ki = F32 ? (uint64_t)pa_inten_key32((float)inp[i]) : pa_inten_key((double)inp[i]);
                        int c = 0;
                        for (int j = b0; j < b1; j++) {
                            const uint64_t kj = F32 ? (uint64_t)pa_inten_key32((float)inp[j]) : pa_inten_key((double)inp[j]);
                            c += (kj > ki) || (kj == ki && j < i);
                        }
                        cnt = c < n_top ? c : 255;
                    }
                }
            }
            const bool keep = cnt < n_top;
            const unsigned bal = __ballot_sync(PA_FULL, keep);
            if (keep) {
                const int pos = out + __popc(bal & below);
                rpk[pos] = make_float2(mzf, __int_as_float(cnt));
                pa_stsf(S_KEY + 4 * pos, mzf);                          // (the keys are not read again)
            }
            out += __popc(bal);
        }
        if (lane == 0) a.rcount[s] = out;
        // m/z cell index over the retained peaks (consumers: pa_match_rank), as k_bin_topn builds it
        if (out <= PA_RCAP && out > 0) {
            __syncwarp();
            const float cbase = pa_ldsf(S_KEY);
            const float cinv = pa_cell_inv(cbase, pa_ldsf(S_KEY + 4 * (out - 1)));
            pa_sts64(S_CELL + 8 * lane, 0x0101010101010101ull * (unsigned long long)out);
            __syncwarp();
            // cell[c] = first retained peak whose cell is >= c: mark the first peak of every occupied cell (empty =
            // `out`), then take the suffix minimum over the 256 cells -- 8 cells per lane, a shuffle scan across lanes
            for (int j = lane; j < out; j += 32) {
                const int cj = pa_cell(pa_ldsf(S_KEY + 4 * j), cbase, cinv);
                const int cp = j > 0 ? pa_cell(pa_ldsf(S_KEY + 4 * j - 4), cbase, cinv) : -1;
                if (cj != cp) pa_sts8(S_CELL + cj, (uint32_t)j);
            }
            __syncwarp();
            unsigned long long v = pa_lds64(S_CELL + 8 * lane);
            unsigned bb[8];
#pragma unroll
            for (int t = 0; t < 8; t++) bb[t] = (unsigned)(v >> (8 * t)) & 0xffu;
#pragma unroll
            for (int t = 6; t >= 0; t--) bb[t] = min(bb[t], bb[t + 1]);
            unsigned x = bb[0];                      // suffix minimum over this and the higher lanes
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_down_sync(PA_FULL, x, o);
                if (lane + o < 32) x = min(x, y);
            }
            unsigned carry = __shfl_down_sync(PA_FULL, x, 1);
            if (lane == 31) carry = (unsigned)out;
            v = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) v |= (unsigned long long)min(bb[t], carry) << (8 * t);
            ((unsigned long long*)(a.ctab + (size_t)s * PA_NCELL))[lane] = v;
            if (lane == 0) a.chead[s] = make_float2(cbase, cinv);
        } else if (lane == 0) a.chead[s] = make_float2(0.f, 0.f);
        __syncwarp();
    }
}
