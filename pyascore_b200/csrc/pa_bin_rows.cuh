// K1 "row form": BinnedSpectra for the spectra that are the rule -- m/z ascending, at most PA_ROWS_MAXCAP peaks (fewer than
// 512 in any one bin), a sane m/z range, n_top <= 31.  cpp/Spectra.cpp:43-68 (bounds + bin index), :24-41 (top n_top of every bin by intensity).
//
// One warp per spectrum, two passes of one peak per lane and 32 peaks per round, nothing staged.  A round of the binning
// pass reads its 32 m/z and intensities straight from global memory (coalesced 8- or 4-byte loads, the next round's already
// in flight), and leaves (float)m/z, the ranking key (float)intensity and the bin in shared memory, plus every bin's
// contiguous run [start, end).  A round of the ranking pass counts, for each of its peaks, the keys of the run that beat
// it (four keys per LDS.128) and writes the kept peaks out at once: one ballot compacts them, and each notes its position
// in the m/z cell index with a shared-memory minimum.  Shared memory is addressed with 32-bit shared-window addresses
// through ld/st.shared (the generic-pointer form recomputes the window base at every access), full rounds carry no bounds
// predicates (the last, partial round is a second instantiation of the same body), and "key > mine" costs 1.5
// instructions (pa_gt_bits).
//
// Everything unusual is DECLINED, not handled: the spectrum's index goes to a list and k_bin_topn (its exact general
// paths included) runs over that list afterwards.  Declined: more peaks than the slot holds, bins or (float)m/z not ascending
// (or NaN), ends outside [0, 1e6], more than PA_NBIN_SMEM bins, and -- in the instantiation for host-narrowed m/z -- the
// spectra that kept an exact float64 copy -- and, found out only at the end, spectra in which two peaks of a bin share a
// float ranking key (k_bin_topn re-ranks such bins on the exact keys).
#pragma once
#include <type_traits>

#define PA_ROWS_MAXCAP 1024
#ifndef PA_K1_MINBLOCKS
#define PA_K1_MINBLOCKS 4      // 64 registers: 32 warps per SM (3 -> 80 registers, 5 -> 48: measured)
#endif
// per warp: key f32[cap + 4] | mzf f32[cap] | bin u8[cap] | range u32[132] | cell u32[256]
#define PA_ROWS_RANGE_BYTES ((PA_NBIN_SMEM + 4) * 4)
#define PA_ROWS_SLOT_BYTES(cap) ((size_t)(cap) * 9 + 16 + PA_ROWS_RANGE_BYTES + PA_NCELL * 4)

// Counting "key > mine" at 1.5 instructions per key: FSET.BF leaves the BITS of 1.0f (0x3f800000 = 127 << 23) or 0, and the
// integer sum of n such words is n * 127 << 23 modulo 2^32, from which n < 512 comes back as ((sum >> 23) * 383) & 511
// (383 is the inverse of 127 modulo 512).  Four keys cost four FSET.BF and two three-input integer adds.
__device__ __forceinline__ int pa_gt_bits(float a, float b) {
    float d;
    asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return __float_as_int(d);
}
#define PA_GT4(v, hi) (pa_gt_bits((v).x, hi) + pa_gt_bits((v).y, hi) + pa_gt_bits((v).z, hi) + pa_gt_bits((v).w, hi))
// the same for the first / last group of a run: key j counts where bit (sh + j) of w is set
#define PA_GT4_IF(v, hi, w, sh)                                                                                   \
    ((((w) >> (sh)) & 1u ? pa_gt_bits((v).x, hi) : 0) + (((w) >> ((sh) + 1)) & 1u ? pa_gt_bits((v).y, hi) : 0) + \
     (((w) >> ((sh) + 2)) & 1u ? pa_gt_bits((v).z, hi) : 0) + (((w) >> ((sh) + 3)) & 1u ? pa_gt_bits((v).w, hi) : 0))
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
__device__ __forceinline__ void pa_atoms_min(uint32_t a, uint32_t v) { asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void pa_sts128(uint32_t a, uint32_t v) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(v)); }

__device__ __forceinline__ void pa_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// F32: float32 intensities (pa_batch.inten32).  NARROW: float32 m/z from the host's narrowing pass (pa_narrow_mz).
template <bool F32, bool NARROW>
__global__ void __launch_bounds__(256, PA_K1_MINBLOCKS) k_bin_rows(PaBinArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename std::conditional<NARROW, float, double>::type mz_t;
    typedef typename std::conditional<F32, float, double>::type in_t;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int cap = a.cap;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw) + (uint32_t)wib * (uint32_t)PA_ROWS_SLOT_BYTES(cap);
    const uint32_t S_KEY = sb;                                  // ranking keys
    const uint32_t S_MZF = sb + (uint32_t)cap * 4 + 16;
    const uint32_t S_BIN = sb + (uint32_t)cap * 8 + 16;
    const uint32_t S_RANGE = S_BIN + (uint32_t)cap + 4;         // start | end << 16 of every bin's run (entry -1: a dummy), then its walk word
    const uint32_t S_CELL = S_RANGE - 4 + PA_ROWS_RANGE_BYTES;
    const int64_t gw = (int64_t)blockIdx.x * wpb + wib, nw = (int64_t)gridDim.x * wpb;
    const int n_top = a.n_top;
    const unsigned below = (1u << lane) - 1u;
    const double dbs = (double)a.bin_size;
    const double inv_bs = __ddiv_rn(1.0, dbs);

    int64_t o0 = 0, o1 = 0;
    if (gw < a.n_spec) { o0 = a.spec_off[gw]; o1 = a.spec_off[gw + 1]; }
    for (int64_t s = gw; s < a.n_spec; s += nw) {
        const int64_t off = o0 - a.peak_base, Pl = o1 - o0;
        if (s + nw < a.n_spec) {
            // the next spectrum's offsets, early -- and its peaks on their way into L2: one prefetch per 128-byte line, the
            // warp covers a whole array with one instruction, so that the loads of the binning pass find them there
            o0 = a.spec_off[s + nw]; o1 = a.spec_off[s + nw + 1];
            const int64_t pn = o1 - o0;
            if (pn > 0) {
                const char* pm = (const char*)((NARROW ? (const mz_t*)a.mz32 : (const mz_t*)a.mz) + (o0 - a.peak_base));
                const char* pi = (const char*)((F32 ? (const in_t*)a.inten32 : (const in_t*)a.inten) + (o0 - a.peak_base));
                const int64_t bm = pn * (int64_t)sizeof(mz_t), bi = pn * (int64_t)sizeof(in_t), at = (int64_t)lane * 128;
                if (at < bm + 128) pa_prefetch_l2(pm + (at < bm ? at : bm - 1));
                if (at < bi + 128) pa_prefetch_l2(pi + (at < bi ? at : bi - 1));
            }
        }
        if (Pl <= 0) { if (lane == 0) { a.rcount[s] = 0; a.chead[s] = make_float2(0.f, 0.f); } continue; }
        auto decline = [&]() { if (lane == 0) a.list[atomicAdd(a.list_n, 1u)] = (int32_t)s; };
        if (Pl > cap) { decline(); continue; }
        if (NARROW) { if (a.esc_off[s] >= 0) { decline(); continue; } }
        const int P = (int)Pl;
        const mz_t* __restrict__ mzp = (NARROW ? (const mz_t*)a.mz32 : (const mz_t*)a.mz) + off;
        const in_t* __restrict__ inp = (F32 ? (const in_t*)a.inten32 : (const in_t*)a.inten) + off;
        // the loads of the first two rounds (the binning pass keeps two rounds in flight ahead of the one it works on)
        mz_t mA = 0, mB = 0, mC = 0;
        in_t kA = 0, kB = 0, kC = 0;
        if (lane < P) { mA = mzp[lane]; kA = inp[lane]; }
        if (lane + 32 < P) { mB = mzp[lane + 32]; kB = inp[lane + 32]; }
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
        pa_sts128z(S_RANGE - 4 + 16 * lane);          // empty bins: start = end = 0
        if (lane == 0) pa_sts128z(S_RANGE - 4 + 512);
        __syncwarp();

        // ---- binning ----
        // The pass relies on what it checks: bins and (float)m/z never decrease from one peak to the next (then every bin
        // is one contiguous run and the output is in m/z order).
        bool sorted = true;
        int carry_bin = -1;                           // "no bin yet": its end lands in the dummy entry before S_RANGE
        float carry_mz = __int_as_float(0xff800000);
        int bq = 0;
        auto row = [&](auto tail, int i, mz_t m, in_t kin) {
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
            const float mzf = NARROW ? (float)m : __double2float_rn((double)m);
            if (valid) {
                pa_stsf(S_MZF + 4 * i, mzf);
                pa_stsf(S_KEY + 4 * i, F32 ? (float)kin : __double2float_rn((double)kin));
                pa_sts8(S_BIN + i, (uint32_t)bq);
            }
            int bprev = __shfl_up_sync(PA_FULL, bq, 1);
            float fprev = __shfl_up_sync(PA_FULL, mzf, 1);
            if (lane == 0) { bprev = carry_bin; fprev = carry_mz; }
            if (valid) {
                sorted = sorted && (bq >= bprev) && (mzf >= fprev);      // NaN fails
                if (bq != bprev) {
                    pa_sts16(S_RANGE + 4 * bq, (uint32_t)i);
                    pa_sts16(S_RANGE + 4 * bprev + 2, (uint32_t)i);
                }
            }
            carry_bin = __shfl_sync(PA_FULL, bq, 31);
            carry_mz = __shfl_sync(PA_FULL, mzf, 31);
        };
        for (int base = 0;;) {
            // A, B, C take turns holding round `base`; the loads of round base + 64 go out before it is worked on
            int in = base + 64 + lane;
            if (in < P) { mC = mzp[in]; kC = inp[in]; }
            if (base + 32 <= P) row(std::false_type(), base + lane, mA, kA); else row(std::true_type(), base + lane, mA, kA);
            base += 32;
            if (base >= P) break;
            in = base + 64 + lane;
            if (in < P) { mA = mzp[in]; kA = inp[in]; }
            if (base + 32 <= P) row(std::false_type(), base + lane, mB, kB); else row(std::true_type(), base + lane, mB, kB);
            base += 32;
            if (base >= P) break;
            in = base + 64 + lane;
            if (in < P) { mB = mzp[in]; kB = inp[in]; }
            if (base + 32 <= P) row(std::false_type(), base + lane, mC, kC); else row(std::true_type(), base + lane, mC, kC);
            base += 32;
            if (base >= P) break;
        }
        if (lane == ((P - 1) & 31)) pa_sts16(S_RANGE + 4 * bq + 2, (uint32_t)P);       // the last run ends with the spectrum
        if (!__all_sync(PA_FULL, sorted)) { decline(); __syncwarp(); continue; }
        __syncwarp();
        // Every run's [start, end) becomes, in place, what a peak of the run needs to walk it in groups of four keys:
        // low half = byte offset of the first group | which of its four keys belong to the run, high half = the same
        // for the last group (no keys when the run sits inside one group: the walk then adds nothing twice).
        int want = 0;                                 // sum over the runs of 0 + 1 + ... + (n - 1): see below
        for (int b = lane; b < n_bins; b += 32) {
            const uint32_t rg = pa_lds32(S_RANGE + 4 * b);
            const uint32_t b0 = rg & 0xffffu, bl = (rg >> 16) - 1u;          // first and last peak of the run
            const int n = (int)(rg >> 16) - (int)b0;
            want += n < 512 ? (n * (n - 1)) >> 1 : -1000000;                 // (pa_gt_count counts up to 511: a larger run declines)
            const uint32_t a0 = b0 & ~3u, al = bl & ~3u;
            uint32_t mf = (0xfu << (b0 & 3u)) & 0xfu;
            uint32_t ml = 0xfu >> (3u - (bl & 3u));
            if (a0 == al) { mf &= ml; ml = 0u; }
            pa_sts32(S_RANGE + 4 * b, (a0 << 2) | mf | (((al << 2) | ml) << 16));
        }
        __syncwarp();

        // ---- ranking + output ----
        // Rank = peaks of the same bin that beat this one, counted on the float keys.  The round's kept peaks go out at
        // once -- {mz, rank} in m/z order, and their positions into the m/z cell index.
        // The cell index (consumers: pa_match_rank) needs a base and a power-of-two cell width that put every kept peak
        // in [0, PA_NCELL): the spectrum's own ends serve, so a peak's cell is known the moment the peak is kept.
        const float cbase = NARROW ? (float)mzp[0] : __double2float_rn(mn);
        const float cinv = pa_cell_inv(cbase, NARROW ? (float)mzp[P - 1] : __double2float_rn(mx));
        pa_sts128(S_CELL + 32 * lane, 0x7fffffffu);
        pa_sts128(S_CELL + 32 * lane + 16, 0x7fffffffu);
        __syncwarp();
        float2* __restrict__ rpk = a.rpk + off;
        int out = 0;
        auto emit = [&](int i, int cnt) {               // (every lane calls; cnt = 255: not kept)
            const bool keep = cnt < n_top;
            const unsigned bal = __ballot_sync(PA_FULL, keep);
            if (keep) {
                const int pos = out + __popc(bal & below);
                const float mzf = pa_ldsf(S_MZF + 4 * i);
                rpk[pos] = make_float2(mzf, __int_as_float(cnt));
                pa_atoms_min(S_CELL + 4 * pa_cell(mzf, cbase, cinv), (uint32_t)pos);    // first kept peak of the cell
            }
            out += __popc(bal);
        };
        int got = 0;
        auto rank = [&](auto tail, int i) {
            constexpr bool TAIL = decltype(tail)::value;
            int c = 255;
            if (!TAIL || i < P) {
                const uint32_t bqi = pa_lds8(S_BIN + i);
                const float hi = pa_ldsf(S_KEY + 4 * i);
                const uint32_t w = pa_lds32(S_RANGE + 4 * bqi);
                uint32_t ga = S_KEY + (w & 0xfff0u);
                const uint32_t gend = S_KEY + ((w >> 16) & 0xfff0u);
                float4 v = pa_lds128f(ga);
                int acc = PA_GT4_IF(v, hi, w, 0);
#pragma unroll 1
                for (ga += 16; ga < gend; ga += 16) {
                    v = pa_lds128f(ga);
                    acc += PA_GT4(v, hi);
                }
                v = pa_lds128f(gend);
                acc += PA_GT4_IF(v, hi, w, 16);
                c = pa_gt_count(acc);
                got += c;
            }
            emit(i, c);
        };
        {
            int base = 0;
            for (; base + 32 <= P; base += 32) rank(std::false_type(), base + lane);
            if (base < P) rank(std::true_type(), base + lane);
        }
        // Ties.  The count of a peak is its rank only if no other peak of its bin has the same float key (equal
        // intensities, +-0, NaN).  Distinct keys give a bin of n peaks the counts 0 .. n-1 in some order; equal keys give
        // some peak a smaller count and none a larger one.  So the counts of the whole spectrum add up to the sum of
        // n (n - 1) / 2 over its bins exactly when no bin holds a tie -- one add per round and two warp reductions --
        // and a spectrum that fails goes to k_bin_topn, which ranks on the exact keys and overwrites what was written here.
        if (__reduce_add_sync(PA_FULL, got) != __reduce_add_sync(PA_FULL, want)) { decline(); __syncwarp(); continue; }
        __syncwarp();
        if (lane == 0) a.rcount[s] = out;
        if (out <= PA_RCAP && out > 0) {
            // cell[c] = first kept peak whose cell is >= c: the suffix minimum over the 256 cells (empty ones hold a
            // large value) -- 8 cells per lane, a shuffle scan across lanes, `out` past the last occupied cell
            const uint4 c0 = pa_lds128u(S_CELL + 32 * lane), c1 = pa_lds128u(S_CELL + 32 * lane + 16);
            unsigned bb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int t = 6; t >= 0; t--) bb[t] = min(bb[t], bb[t + 1]);
            unsigned x = bb[0];                      // suffix minimum over this and the higher lanes
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_down_sync(PA_FULL, x, o);
                if (lane + o < 32) x = min(x, y);
            }
            unsigned carry = __shfl_down_sync(PA_FULL, x, 1);
            if (lane == 31) carry = (unsigned)out;
            carry = min(carry, (unsigned)out);
            unsigned long long v = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) v |= (unsigned long long)min(bb[t], carry) << (8 * t);
            ((unsigned long long*)(a.ctab + (size_t)s * PA_NCELL))[lane] = v;
            if (lane == 0) a.chead[s] = make_float2(cbase, cinv);
        } else if (lane == 0) a.chead[s] = make_float2(0.f, 0.f);
        __syncwarp();
    }
}
