// K3b, second form: Ascores at the granularity of ONE site-determining-ion merge.
//
// cpp/Ascore.cpp:212-254 (calculateAscores), :157-210 (calculateAmbiguity), cpp/ModifiedPeptide.cpp:259-320
// (getSiteDeterminingIons).  The unit of work of the first form was a (PSM, modified site) entry: one thread ran
// every tied competitor and every ion type of its entry, so a kernel lasted as long as its longest entry (config 5:
// 7 % of the warp slots busy).  Here a warp takes 32 entries of the work-sorted list and spreads their
// (entry, tied competitor, ion type) merges -- the ITEMS -- over its lanes, round after round; the two ion types of a
// (entry, competitor) pair sit in neighbouring lanes and are summed with one shuffle, the minimum over an entry's
// competitors is taken in shared memory.
//
// One merge, per lane:
//  * the float32 running sums of the walk (the reference's sequential adds, cpp/ModifiedPeptide.cpp:379-392) are
//    built ONCE per list into shared memory, [step][thread]; every (neutral-loss sum, charge) stream of the list is
//    then a cursor over that sequence, and advancing a stream is one shared-memory load + the FP64 m/z formula
//    instead of a peptide byte, a mass table and a loss table fetched from global memory per stream and step;
//  * common prefix: every fragment below T stems from steps both isoforms share, where T bounds from below every
//    fragment of the steps from the first differing residue on; the sorted lists therefore start with the same values,
//    which the greedy merge drops pairwise (|x - y| = 0 < mz_error): the streams start at their first fragment >= T;
//  * common suffix: once the walk has passed both moved sites and the running sums (and neutral-loss states) of the
//    two isoforms agree bit for bit at some step i0, they agree ever after.  When every stream of both lists is past
//    i0, no popped head is pending and both lists have the same number of elements left, those elements are the same
//    multiset (upper sets of equal size of one sorted multiset) and drop pairwise: the merge stops there.
// Both cuts only remove pairs the reference's merge would drop, so hits and trials are the reference's.
#pragma once

#define ASC_LC 32                 // walk steps held in shared memory per list (peptides up to 33 residues); longer
                                  // peptides and unsupported shapes go to k_ascore_generic

template <int NQ, bool NL>
struct AscSm2 {
    float seq[2][ASC_LC][ASC_BLOCK];                         // running sums of lists A (best isoform) and B (competitor)
    uint8_t nls[NL ? 2 : 1][NL ? ASC_LC : 1][ASC_BLOCK];     // neutral-loss state after each step
    float val[NQ > 4 ? 2 : 1][NQ > 4 ? NQ : 1][ASC_BLOCK];   // pending fragment of each stream (+inf: exhausted)
    uint8_t stp[NQ > 4 ? 2 : 1][NQ > 4 ? NQ : 1][ASC_BLOCK]; // step of each stream's pending fragment
    int off[ASC_BLOCK / 32][33];                             // per warp: first item of each entry, [32] = total
    float asc[ASC_BLOCK / 32][32];                           // per warp: running minimum of each entry
    int gen[ASC_BLOCK / 32][32];                             // per warp: entry needs the generic kernel
};

// fragment m/z of running sum `run` minus neutral-loss sum `sig` at charge z (cpp/ModifiedPeptide.cpp:570-591)
__device__ __forceinline__ float asc_frag(float run, float sig, bool nl, double a1, double a2, int z) {
    const float base = nl ? __fsub_rn(run, sig) : run;
    return pa_charge_mz(__dsub_rn(__dadd_rn((double)base, a1), a2), z);
}

// One merge: ion type `type` of the best isoform (residue mask alo/ahi) against the competitor (blo/bhi).
// WARP-SYNCHRONOUS: every lane of the warp calls it in the same round (lanes without an item pass live = false) and
// the merge loop votes once per trip, so that the lanes of a warp stay on the same instruction -- left to the
// scheduler, the lanes of such a long, data-dependent loop drift apart for good (1 of 32 lanes per instruction
// in the first capture of this kernel, profiles/r03d).  Returns false when the shape is not covered (-> generic kernel).
template <int NQ, bool NL>
__device__ __forceinline__ bool asc_merge2(bool live, const PaCfg& cfg, const AscPep& q, const AscPeaks& pk,
                                           AscSm2<NQ, NL>* sm, const float* s_nl, char type, uint64_t alo, uint64_t ahi,
                                           uint64_t blo, uint64_t bhi, int depth, int& hitsA, int& trialsA, int& hitsB,
                                           int& trialsB) {
    const int tid = threadIdx.x;
    const int L = q.L, Z = q.Z, steps = L - 1;
    const float PINF = __int_as_float(0x7f800000);
    bool ok = true;
    if (live && (L < 2 || steps > ASC_LC)) { live = false; ok = false; }
    const bool fwd = (type == 'b' || type == 'c');
    double a1, a2;
    pa_type_consts(type, a1, a2);
    const uint8_t* nvar = cfg.nl_nvar;
    // first / last walk step whose residue differs between the two isoforms
    int d = 0, dl = 0;
    if (live) {
        const uint64_t xlo = alo ^ blo, xhi = ahi ^ bhi;
        const int p_first = xlo ? __ffsll((long long)xlo) - 1 : (xhi ? 63 + __ffsll((long long)xhi) : 128);
        const int p_last = xhi ? 127 - __clzll((long long)xhi) : (xlo ? 63 - __clzll((long long)xlo) : -1);
        d = fwd ? p_first : L - 1 - p_last;
        dl = fwd ? p_last : L - 1 - p_first;
        d = d < 0 ? 0 : d;
        if (d >= steps) live = false;        // every walked residue is common: all fragments drop pairwise
    }
    // ---- the two running-sum sequences (B shares A's steps before d) ----
    int finA = 0, finB = 0;
    if (live) {
        bool mono = true;
        float runA = 0.f, runB = 0.f;
        int nlsA = 0, nlsB = 0;
        for (int s = 0; s < steps; s++) {
            const int i = fwd ? s : L - 1 - s;
            int idxA, idxB;
            const int stA = (int)(((i < 64) ? (alo >> i) : (ahi >> (i - 64))) & 1ull);
            const int stB = (int)(((i < 64) ? (blo >> i) : (bhi >> (i - 64))) & 1ull);
            const float rA = asc_res(cfg, q, i, stA, idxA);
            float rB = rA;
            idxB = idxA;
            if (stB != stA) rB = asc_res(cfg, q, i, stB, idxB);
            const float nA = (s == 0) ? rA : __fadd_rn(rA, runA);
            const float nB = (s == 0) ? rB : __fadd_rn(rB, runB);
            if (s > 0 && (nA < runA || nB < runB)) mono = false;
            runA = nA; runB = nB;
            sm->seq[0][s][tid] = nA; sm->seq[1][s][tid] = nB;
            if (NL) {
                if (idxA) nlsA = pa_nl_bump(nlsA, idxA);
                if (idxB) nlsB = pa_nl_bump(nlsB, idxB);
                sm->nls[0][s][tid] = (uint8_t)nlsA; sm->nls[1][s][tid] = (uint8_t)nlsB;
            }
        }
        finA = nlsA; finB = nlsB;
        if (!mono) { live = false; ok = false; }          // a non-positive residue mass: the streams are not ascending
    }
    __syncwarp();
    const int VA = NL ? nvar[finA] : 1, VB = NL ? nvar[finB] : 1;
    if (live && (VA * Z > NQ || VB * Z > NQ)) { live = false; ok = false; }
    // ---- suffix: first step after both moved sites at which the two walks agree bit for bit ----
    int i0 = steps;
    // ---- prefix: T = lower bound of every fragment of the steps >= d (largest loss sum at step d, every charge) ----
    float T = -PINF;
    if (live && cfg.err > 0.f) {
        for (int s = dl + 1; s < steps; s++) {
            const bool same = __float_as_int(sm->seq[0][s][tid]) == __float_as_int(sm->seq[1][s][tid]) &&
                              (!NL || sm->nls[0][s][tid] == sm->nls[1][s][tid]);
            if (same && i0 == steps) i0 = s;
        }
        if (d > 0) {
            T = PINF;
            const float sgA = NL ? s_nl[finA * 16 + VA - 1] : 0.f, sgB = NL ? s_nl[finB * 16 + VB - 1] : 0.f;
            const float rdA = sm->seq[0][d][tid], rdB = sm->seq[1][d][tid];
            for (int z = 1; z <= Z; z++) T = fminf(T, fminf(asc_frag(rdA, sgA, NL, a1, a2, z), asc_frag(rdB, sgB, NL, a1, a2, z)));
        }
    }
    __syncwarp();
    // ---- streams: q = v * Z + (z - 1); state in registers up to four streams per list, else in shared memory ----
    float valA[NQ > 4 ? 1 : NQ], valB[NQ > 4 ? 1 : NQ];
    int stpA[NQ > 4 ? 1 : NQ], stpB[NQ > 4 ? 1 : NQ];
    int leftA = 0, leftB = 0, lagA = 0, lagB = 0;      // elements left / streams still before i0, per list
    // first pending fragment of stream qi of list w: its first fragment >= T (walking down from step d: a stream is
    // non-decreasing along the walk), not before the step at which its loss sum first exists
    auto stream_start = [&](int w, int qi, int V, int fin, float& v0, int& s0) {
        v0 = PINF; s0 = steps;
        if (!live || qi >= V * Z) return;
        const int v = NL ? qi / Z : 0, z = qi - v * Z + 1;
        const float sg = NL ? s_nl[fin * 16 + v] : 0.f;
        int av = 0;
        if (NL && v > 0) {                   // (the stack only grows: once available, a sum stays available)
            av = steps;
            int prev = -1;
            for (int s = 0; s < steps; s++) {
                const int st = sm->nls[w][s][tid];
                if (st != prev && av == steps) {
                    bool in = false;
                    const int nv = nvar[st];
                    for (int u = 0; u < nv; u++) in |= (s_nl[st * 16 + u] == sg);
                    if (in) av = s;
                }
                prev = st;
            }
        }
        s0 = av > d ? av : d;
        if (s0 >= steps) { s0 = steps; return; }
        v0 = asc_frag(sm->seq[w][s0][tid], sg, NL, a1, a2, z);
        for (int sd = s0 - 1; sd >= av; sd--) {
            const float vp = asc_frag(sm->seq[w][sd][tid], sg, NL, a1, a2, z);
            if (!(vp >= T)) break;
            v0 = vp; s0 = sd;
        }
    };
    if (NQ > 4) {
        for (int w = 0; w < 2; w++) {
            const int V = w ? VB : VA, fin = w ? finB : finA;
            int left = 0, lag = 0;
            for (int qi = 0; qi < NQ; qi++) {
                float v0; int s0;
                stream_start(w, qi, V, fin, v0, s0);
                sm->val[w][qi][tid] = v0; sm->stp[w][qi][tid] = (uint8_t)s0;
                left += steps - s0; lag += s0 < i0;
            }
            if (w) { leftB = left; lagB = lag; } else { leftA = left; lagA = lag; }
        }
    } else {
#pragma unroll
        for (int qi = 0; qi < (NQ > 4 ? 1 : NQ); qi++) {
            stream_start(0, qi, VA, finA, valA[qi], stpA[qi]);
            stream_start(1, qi, VB, finB, valB[qi], stpB[qi]);
            leftA += steps - stpA[qi]; lagA += stpA[qi] < i0;
            leftB += steps - stpB[qi]; lagB += stpB[qi] < i0;
        }
    }
    if (!live) { leftA = 0; leftB = 0; }
    __syncwarp();
    // ---- greedy tolerance merge (cpp/ModifiedPeptide.cpp:288-316) ----
    float x = 0.f, y = 0.f;
    bool hx = false, hy = false;             // a popped element is pending
    // pop the smallest pending fragment of list w and advance its stream
    auto pop = [&](int w, float& out) {
        int bq = 0, sb;
        float xm;
        if (NQ > 4) {
            const int nq = (w ? VB : VA) * Z;
            xm = sm->val[w][0][tid];
            for (int i = 1; i < nq; i++) { const float v = sm->val[w][i][tid]; if (v < xm) { xm = v; bq = i; } }
            sb = sm->stp[w][bq][tid];
        } else {
            xm = w ? valB[0] : valA[0]; sb = w ? stpB[0] : stpA[0];
#pragma unroll
            for (int i = 1; i < (NQ > 4 ? 1 : NQ); i++) {
                const float v = w ? valB[i] : valA[i];
                if (v < xm) { xm = v; bq = i; sb = w ? stpB[i] : stpA[i]; }
            }
        }
        const int step = sb + 1;
        float nv = PINF;
        if (step < steps) {
            int v = 0, z = bq + 1;
            if (NL) { v = (Z == 1) ? bq : (Z == 2 ? bq >> 1 : bq / Z); z = bq - v * Z + 1; }
            const float sg = NL ? s_nl[(w ? finB : finA) * 16 + v] : 0.f;
            nv = asc_frag(sm->seq[w][step][tid], sg, NL, a1, a2, z);
        }
        if (NQ > 4) { sm->val[w][bq][tid] = nv; sm->stp[w][bq][tid] = (uint8_t)step; }
        else {
#pragma unroll
            for (int i = 0; i < (NQ > 4 ? 1 : NQ); i++)
                if (i == bq) { if (w) { valB[i] = nv; stpB[i] = step; } else { valA[i] = nv; stpA[i] = step; } }
        }
        if (step == i0) { if (w) lagB--; else lagA--; }
        out = xm;
    };
    bool more = live && (leftA > 0 || leftB > 0);
    while (__any_sync(PA_FULL, more)) {
        if (more) {
            if (!hx && leftA > 0) { pop(0, x); hx = true; leftA--; }
            if (!hy && leftB > 0) { pop(1, y); hy = true; leftB--; }
            int takeA;                           // 1: A survives, 0: B survives, -1: both dropped
            if (!hy) takeA = 1;
            else if (!hx) takeA = 0;
            else if (fabsf(__fsub_rn(x, y)) < cfg.err) takeA = -1;
            else takeA = x < y;
            if (takeA < 0) {
                hx = false; hy = false;
                // both lists are past the last differing fragment and equally long: the rest drops pairwise
                if ((lagA | lagB) == 0 && leftA == leftB) { leftA = 0; leftB = 0; }
            } else {
                asc_survivor(cfg, pk, depth, takeA ? x : y, takeA != 0, hitsA, trialsA, hitsB, trialsB);
                if (takeA) hx = false; else hy = false;
            }
            more = hx || hy || leftA > 0 || leftB > 0;
        }
    }
    return ok;
}

// One item = one tied competitor of one (PSM, modified site) entry, ion types [type_lo, type_hi).  What the item
// needs to know about its entry is rebuilt from global memory (L1-resident: the items of an entry sit in
// neighbouring lanes / rounds).  Warp-synchronous like asc_merge2: lanes without an item pass active = false.
template <int NQ, bool NL>
__device__ __forceinline__ bool asc_item(bool active, const PaCfg& cfg, const PaBatchDev& b, const PaAscArgs& a,
                                         AscSm2<NQ, NL>* sm, const float* s_nl, int64_t t, int comp, int type_lo,
                                         int type_hi, int& depth, int& hitsA, int& trialsA, int& hitsB, int& trialsB,
                                         bool& trivial) {
    hitsA = hitsB = trialsA = trialsB = 0;
    depth = 0;
    trivial = false;
    AscPep q;
    q.pep = b.pep; q.L = 0; q.Z = 1; q.a0 = 0; q.a1 = 0; q.aux_pos = b.aux_pos; q.aux_mass = b.aux_mass; q.aux_lo = 0; q.aux_hi = 0;
    AscPeaks pk;
    pk.pp = b.rpk; pk.R = 0; pk.ctab = b.ctab; pk.cbase = 0.f; pk.cinv = 0.f;
    uint64_t alo = 0, ahi = 0, blo = 0, bhi = 0;
    bool ok = true, live = active;
    if (active) {
        const int32_t p = a.mod_psm[t];
#ifdef PA_DEBUG_ITEMS
        if (t < 0 || t >= a.n_entries || p < 0 || a.best_idx[p] == 0xffffffffu)
            printf("asc_item: bad item t=%lld p=%d comp=%d n_entries=%lld lane=%d blk=%d\n", (long long)t, p, comp,
                   (long long)a.n_entries, threadIdx.x, blockIdx.x);
#endif
        const int j = (int)(a.mod_lo + t - a.mod_off[p]);
        const int S = a.psm_S[p], k = b.n_mod[p];
        const int64_t ib = a.iso_off[p];
        const int po = b.pep_off[p];
        q.pep = b.pep + po; q.L = b.pep_off[p + 1] - po; q.Z = b.max_charge[p];
        if (b.aux_off != nullptr) {
            q.a0 = b.aux_off[p]; q.a1 = b.aux_off[p + 1];
            for (int x = q.a0; x < q.a1; x++) {
                const uint32_t pos = q.aux_pos[x];
                const int idx = pos > 0 ? (int)pos - 1 : 0;
                if (idx < 64) q.aux_lo |= 1ull << idx; else if (idx < 128) q.aux_hi |= 1ull << (idx - 64);
            }
        }
        if (q.L < 2 || (!NL && (cfg.has_nl || q.Z > NQ))) { ok = false; live = false; }
        else {
            const int sp = b.psm_spec[p];
            pk.pp = b.rpk + (b.spec_off[sp] - b.spec_base); pk.R = b.rcount[sp];
            pk.ctab = b.ctab + (size_t)sp * PA_NCELL;
            { const float2 chead = b.chead[sp]; pk.cbase = chead.x; pk.cinv = chead.y; }
            const uint32_t best = a.best_idx[p];
            const uint64_t best_bits = pa_unrank(cfg.binom, S, k, best);
            uint64_t rem = best_bits;
            for (int jj = 0; jj < j; jj++) rem &= rem - 1;
            const int site = __ffsll((long long)rem) - 1;
            // competitor: the mod moves from `site` to the comp-th site of the tie set
            uint64_t tt = a.tie[t];
            for (int c = 0; c < comp; c++) tt &= tt - 1;
            const int u = __ffsll((long long)tt) - 1;
            // residue masks of both isoforms
            int pos_site = 0, pos_u = 0;
            {
                int n = 0;
                for (int i = 0; i < q.L && n < 64; i++) {
                    const int c = (int)q.pep[i] - 'A';
                    const bool is = ((cfg.mod_letters >> c) & 1u) || (cfg.allow_n && i == 0) || (cfg.allow_c && i == q.L - 1);
                    if (!is) continue;
                    if ((best_bits >> n) & 1ull) { if (i < 64) alo |= 1ull << i; else ahi |= 1ull << (i - 64); }
                    if (n == site) pos_site = i;
                    if (n == u) pos_u = i;
                    n++;
                }
            }
            blo = alo; bhi = ahi;
            if (pos_site < 64) blo &= ~(1ull << pos_site); else bhi &= ~(1ull << (pos_site - 64));
            if (pos_u < 64) blo |= 1ull << pos_u; else bhi |= 1ull << (pos_u - 64);
            const uint32_t ci = pa_rank(cfg.binom, S, k, (best_bits & ~(1ull << site)) | (1ull << u));
            const float wA = a.iso.w[ib + best], wB = a.iso.w[ib + ci];
            if ((double)fabsf(__fsub_rn(wA, wB)) < 1e-6) { trivial = true; live = false; }     // cpp/Ascore.cpp:161-163
            else {
                // depth with the largest score difference (cpp/Ascore.cpp:165-175), first strict maximum
                const unsigned long long loA = a.iso.lo[ib + best], hiA = a.iso.hi[ib + best];
                const unsigned long long loB = a.iso.lo[ib + ci], hiB = a.iso.hi[ib + ci];
                const int nfA = (int)a.iso.nfrag[ib + best], nfB = (int)a.iso.nfrag[ib + ci];
                float max_diff = 0.f;
                for (int dd = 0; dd < PA_N_TOP; dd++) {
                    const float diff = __fsub_rn(__ldg(cfg.T + pa_tab_index(nfA, pa_cum_get(loA, hiA, dd), dd)),
                                                 __ldg(cfg.T + pa_tab_index(nfB, pa_cum_get(loB, hiB, dd), dd)));
                    if (diff > max_diff) { max_diff = diff; depth = dd; }
                }
            }
        }
    }
    __syncwarp();
    // the type loop is warp-uniform: one type per lane when the two ion types of a pair sit in neighbouring lanes
    // (type_hi - type_lo == 1 for every lane), else every configured type for every lane
    const int nt = type_hi - type_lo;
    for (int ti = 0; ti < nt; ti++) {
        const bool r = asc_merge2<NQ, NL>(live, cfg, q, pk, sm, s_nl, cfg.types[type_lo + ti], alo, ahi, blo, bhi, depth,
                                          hitsA, trialsA, hitsB, trialsB);
        if (!r) { ok = false; live = false; }
    }
    return ok;
}

// minimum of floats of either sign through integer atomics (the slot starts at +inf)
__device__ __forceinline__ void asc_atomic_min(float* slot, float v) {
    if (v >= 0.f) atomicMin((int*)slot, __float_as_int(v));
    else atomicMax((unsigned int*)slot, __float_as_uint(v));
}

#ifndef PA_ASC2_MINB
#define PA_ASC2_MINB 4
#endif
template <int NQ, int CLS, bool NL>
__global__ void __launch_bounds__(ASC_BLOCK, (NQ > 4 ? 3 : PA_ASC2_MINB)) k_ascore_items(PaCfg cfg, PaBatchDev b, PaAscArgs a) {
    extern __shared__ __align__(16) unsigned char asc2_smem_raw[];
    AscSm2<NQ, NL>* sm = (AscSm2<NQ, NL>*)asc2_smem_raw;
    const float* s_nl = cfg.nl_sums;         // [256][16] loss sums per stack state: a few rows in use, L1-resident
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t n_cls = a.work_count[CLS];
    int64_t first = 0;                       // classes are contiguous in the sorted list
#pragma unroll
    for (int c = 0; c < CLS; c++) first += a.work_count[c];
#ifndef PA_ASC_PAIR
#define PA_ASC_PAIR 0
#endif
    // PA_ASC_PAIR: the two ion types of an (entry, competitor) pair go to neighbouring lanes (twice the items, but a
    // b-walk and a y-walk of one pair have complementary lengths, so a round lasts as long as the longer of the two);
    // default: one lane runs both walks of its pair, whose lengths add up to about the same for every pair
    const bool pair = PA_ASC_PAIR && cfg.n_types == 2;
    const int ty = pair ? 2 : 1;
    int* s_off = sm->off[wib];
    float* s_asc = sm->asc[wib];
    int* s_gen = sm->gen[wib];
    const int64_t warps = (int64_t)gridDim.x * (ASC_BLOCK / 32);
    for (int64_t base = ((int64_t)blockIdx.x * (ASC_BLOCK / 32) + wib) * 32; base < n_cls; base += warps * 32) {
        const bool have = base + lane < n_cls;
        const int64_t t = have ? (int64_t)a.work_sorted[first + base + lane] : -1;
        int n_items = have ? __popcll(a.tie[t]) * ty : 0;
        int off = n_items;
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(PA_FULL, off, o); if (lane >= o) off += v; }
        const int total = __shfl_sync(PA_FULL, off, 31);
        off -= n_items;
        __syncwarp();
        s_off[lane] = off;
        if (lane == 31) s_off[32] = total;
        s_asc[lane] = __int_as_float(0x7f800000);
        s_gen[lane] = 0;
        __syncwarp();
        for (int r0 = 0; r0 < total; r0 += 32) {
            const int r = r0 + lane;
            const bool active = r < total;
            int e = 0;
            if (active) {                    // entry of item r: last lane whose first item is <= r (empty entries skipped)
                int lo = 0, hi = 32;
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= r) lo = mid; else hi = mid; }
                e = lo;
            }
            const int64_t te = __shfl_sync(PA_FULL, t, e);
            int depth = 0, hA = 0, tA = 0, hB = 0, tB = 0;
            bool trivial = false;
            const int jn = active ? r - s_off[e] : 0;
            const int comp = pair ? jn >> 1 : jn, t0 = pair ? (jn & 1) : 0, t1 = pair ? t0 + 1 : cfg.n_types;
            bool ok = asc_item<NQ, NL>(active, cfg, b, a, sm, s_nl, te, comp, t0, t1, depth, hA, tA, hB, tB, trivial);
            if (pair) {                      // items come in (type 0, type 1) lane pairs: r even <-> lane even
                hA += __shfl_xor_sync(PA_FULL, hA, 1); tA += __shfl_xor_sync(PA_FULL, tA, 1);
                hB += __shfl_xor_sync(PA_FULL, hB, 1); tB += __shfl_xor_sync(PA_FULL, tB, 1);
                const int ok2 = __shfl_xor_sync(PA_FULL, (int)ok, 1);
                ok = ok && ok2;
            }
            if (active && (!pair || !(lane & 1))) {
                if (!ok) s_gen[e] = 1;
                else {
                    float amb = 0.f;
                    if (!trivial) amb = __fsub_rn(__ldg(cfg.T + pa_tab_index(tA, hA, depth)), __ldg(cfg.T + pa_tab_index(tB, hB, depth)));
                    asc_atomic_min(&s_asc[e], amb);
                }
            }
        }
        __syncwarp();
        if (have) {
            if (s_gen[lane]) a.generic_list[atomicAdd(a.generic_count, 1)] = (int32_t)t;
            else if (a.ascores) a.ascores[a.mod_lo + t] = s_asc[lane];
        }
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------------------------
// K3b, third form: the unit of work is one (entry, tied competitor) pair, handed out 32 at a time from a global
// cursor to one resident wave of warps.  The pairs are numbered through an exclusive scan over the work-sorted
// entry list (stream class, then estimated merge length, longest first), so a kernel no longer lasts as long as
// its heaviest entry (config 5: entries with up to 9 tied competitors, 6-8 % of the warp slots busy) or its
// heaviest warp.  A pair runs the register / stream merges of the first form (asc_merge_type) for every ion type
// and lowers the entry's Ascore with an atomic minimum; entries a merge cannot cover go, once, to k_ascore_generic.
// ---------------------------------------------------------------------------------------------------------------
struct PaAscItemArgs {
    int32_t* item_cnt;           // [n_entries + 1] tied competitors of the i-th entry of the sorted list (0 past the work)
    const int32_t* item_off;     // [n_entries + 1] exclusive scan of item_cnt
    unsigned long long* cursor;  // [4] next pair of each stream class (zeroed before the launch)
    int* gen_flag;               // [n_entries] entry already queued for the generic kernel (zeroed)
};

// pairs per entry of the sorted work list; Ascores of the entries with work start at +inf
__global__ void k_asc_item_count(PaAscArgs a, PaAscItemArgs it) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > a.n_entries) return;
    const int64_t n_work = (int64_t)a.work_count[0] + a.work_count[1] + a.work_count[2] + a.work_count[3];
    int c = 0;
    if (i < n_work) {
        const int64_t t = a.work_sorted[i];
        c = __popcll(a.tie[t]);
        if (a.ascores) a.ascores[a.mod_lo + t] = __int_as_float(0x7f800000);
    }
    it.item_cnt[i] = c;
}

// Ambiguity of the best isoform of entry t against its comp-th tied competitor (cpp/Ascore.cpp:157-210).
// Returns false when a merge does not cover the shape (-> generic kernel).
template <int NQ>
__device__ __forceinline__ bool asc_pair(const PaCfg& cfg, const PaBatchDev& b, const PaAscArgs& a, int64_t t, int comp,
                                         float& amb) {
    const int32_t p = a.mod_psm[t];
    const int j = (int)(a.mod_lo + t - a.mod_off[p]);
    const int S = a.psm_S[p], k = b.n_mod[p];
    const int64_t ib = a.iso_off[p];
    AscPep q;
    const int po = b.pep_off[p];
    q.pep = b.pep + po; q.L = b.pep_off[p + 1] - po; q.Z = b.max_charge[p];
    q.a0 = 0; q.a1 = 0; q.aux_pos = b.aux_pos; q.aux_mass = b.aux_mass; q.aux_lo = 0; q.aux_hi = 0;
    if (b.aux_off != nullptr) {
        q.a0 = b.aux_off[p]; q.a1 = b.aux_off[p + 1];
        for (int x = q.a0; x < q.a1; x++) {
            const uint32_t pos = q.aux_pos[x];
            const int idx = pos > 0 ? (int)pos - 1 : 0;
            if (idx < 64) q.aux_lo |= 1ull << idx; else if (idx < 128) q.aux_hi |= 1ull << (idx - 64);
        }
    }
    amb = 0.f;
    if ((q.L < 2) || (NQ <= 4 && (cfg.has_nl || q.Z > NQ))) return false;
    const int sp = b.psm_spec[p];
    AscPeaks pk;
    pk.pp = b.rpk + (b.spec_off[sp] - b.spec_base); pk.R = b.rcount[sp];
    pk.ctab = b.ctab + (size_t)sp * PA_NCELL;
    { const float2 chead = b.chead[sp]; pk.cbase = chead.x; pk.cinv = chead.y; }
    const uint32_t best = a.best_idx[p];
    const uint64_t best_bits = pa_unrank(cfg.binom, S, k, best);
    uint64_t rem = best_bits;
    for (int jj = 0; jj < j; jj++) rem &= rem - 1;
    const int site = __ffsll((long long)rem) - 1;
    uint64_t tt = a.tie[t];
    for (int c = 0; c < comp; c++) tt &= tt - 1;
    const int u = __ffsll((long long)tt) - 1;
    // residue masks of both isoforms: the competitor has the mod of `site` on site u instead
    uint64_t alo = 0, ahi = 0;
    int pos_site = 0, pos_u = 0;
    {
        int n = 0;
        for (int i = 0; i < q.L && n < 64; i++) {
            const int c = (int)q.pep[i] - 'A';
            const bool is = ((cfg.mod_letters >> c) & 1u) || (cfg.allow_n && i == 0) || (cfg.allow_c && i == q.L - 1);
            if (!is) continue;
            if ((best_bits >> n) & 1ull) { if (i < 64) alo |= 1ull << i; else ahi |= 1ull << (i - 64); }
            if (n == site) pos_site = i;
            if (n == u) pos_u = i;
            n++;
        }
    }
    uint64_t blo = alo, bhi = ahi;
    if (pos_site < 64) blo &= ~(1ull << pos_site); else bhi &= ~(1ull << (pos_site - 64));
    if (pos_u < 64) blo |= 1ull << pos_u; else bhi |= 1ull << (pos_u - 64);
    const uint32_t ci = pa_rank(cfg.binom, S, k, (best_bits & ~(1ull << site)) | (1ull << u));
    const float wA = a.iso.w[ib + best], wB = a.iso.w[ib + ci];
    if ((double)fabsf(__fsub_rn(wA, wB)) < 1e-6) return true;                 // cpp/Ascore.cpp:161-163: ambiguity 0
    // depth with the largest score difference (cpp/Ascore.cpp:165-175), first strict maximum
    int depth = 0;
    {
        const unsigned long long loA = a.iso.lo[ib + best], hiA = a.iso.hi[ib + best];
        const unsigned long long loB = a.iso.lo[ib + ci], hiB = a.iso.hi[ib + ci];
        const int nfA = (int)a.iso.nfrag[ib + best], nfB = (int)a.iso.nfrag[ib + ci];
        float max_diff = 0.f;
        for (int d = 0; d < PA_N_TOP; d++) {
            const float diff = __fsub_rn(__ldg(cfg.T + pa_tab_index(nfA, pa_cum_get(loA, hiA, d), d)),
                                         __ldg(cfg.T + pa_tab_index(nfB, pa_cum_get(loB, hiB, d), d)));
            if (diff > max_diff) { max_diff = diff; depth = d; }
        }
    }
    int hitsA = 0, hitsB = 0, trialsA = 0, trialsB = 0;
    for (int ti = 0; ti < cfg.n_types; ti++)
        if (!asc_merge_type<NQ>(cfg, q, pk, cfg.types[ti], alo, ahi, blo, bhi, depth, hitsA, trialsA, hitsB, trialsB))
            return false;
    amb = __fsub_rn(__ldg(cfg.T + pa_tab_index(trialsA, hitsA, depth)), __ldg(cfg.T + pa_tab_index(trialsB, hitsB, depth)));
    return true;
}

#ifndef PA_ASC3_GRAB
#define PA_ASC3_GRAB 1           // rounds of 32 pairs a warp takes per visit to the cursor
#endif
template <int NQ, int CLS>
__global__ void __launch_bounds__(128, (NQ <= 1 ? PA_ASC_MINBLOCKS : (NQ <= 2 ? PA_ASC2_MINBLOCKS : (NQ <= 4 ? PA_ASC4_MINBLOCKS : PA_ASC12_MINBLOCKS))))
k_ascore_pairs(PaCfg cfg, PaBatchDev b, PaAscArgs a, PaAscItemArgs it) {
    const int lane = threadIdx.x & 31;
    int64_t first = 0;                       // classes are contiguous in the sorted entry list, hence in the pair numbering
#pragma unroll
    for (int c = 0; c < CLS; c++) first += a.work_count[c];
    const int64_t last = first + a.work_count[CLS];
    const int64_t g0 = it.item_off[first], g1 = it.item_off[last];
    if (g1 <= g0) return;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(it.cursor + CLS, 32ull * PA_ASC3_GRAB);
        base = __shfl_sync(PA_FULL, base, 0);
        if ((int64_t)base >= g1 - g0) break;
        for (int rr = 0; rr < PA_ASC3_GRAB; rr++) {
            const int64_t g = g0 + (int64_t)base + rr * 32 + lane;
            if (g >= g1) continue;
            // entry of pair g: last position of [first, last) whose first pair is <= g
            int64_t lo = first, hi = last;
            while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)it.item_off[mid] <= g) lo = mid; else hi = mid; }
            const int64_t t = a.work_sorted[lo];
            float amb;
            if (asc_pair<NQ>(cfg, b, a, t, (int)(g - it.item_off[lo]), amb)) {
                if (a.ascores) asc_atomic_min(a.ascores + a.mod_lo + t, amb);
            } else if (atomicExch(it.gen_flag + t, 1) == 0)
                a.generic_list[atomicAdd(a.generic_count, 1)] = (int32_t)t;
        }
    }
}
