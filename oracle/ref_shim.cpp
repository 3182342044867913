// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// extern "C" shim over the UNMODIFIED reference C++ sources, which are compiled
// where they lie under /root/reference (see oracle/Makefile; output goes to the
// git-ignored oracle/_ref/).  No reference source is copied into this repo: the
// four .cpp files are pulled in by path, as one translation unit, exactly like
// the reference's own Cython unity build does
// (pyascore/ptm_scoring/Ascore.pxd:6-7, ModifiedPeptide.pxd:6-7, Spectra.pxd,
// Util.pxd: `cdef extern from "cpp/X.cpp"`).
//
// The driver loop in refshim_score() restates what the Cython method
// PyAscore.score does around the C++ classes (pyascore/ptm_scoring/Ascore.pyx:129-152).
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>
#include <string>
#include <vector>

#include "cpp/Util.cpp"
#include "cpp/Spectra.cpp"
#include "cpp/ModifiedPeptide.cpp"
#include "cpp/Ascore.cpp"

using namespace ptmscoring;

struct RefScorer {
    BinnedSpectra spectra;
    ModifiedPeptide peptide;
    Ascore ascore;
    RefScorer(float bin_size, size_t n_top, const char* group, float mass, float err, const char* types)
        : spectra(bin_size, n_top), peptide(group, mass, err, types), ascore() {}
};

static std::vector<ScoreContainer> g_tmp;

extern "C" {

void* refshim_new(float bin_size, size_t n_top, const char* mod_group, float mod_mass,
                  float mz_error, const char* fragment_types) {
    return new RefScorer(bin_size, n_top, mod_group, mod_mass, mz_error, fragment_types);
}

void refshim_free(void* h) { delete (RefScorer*)h; }

void refshim_add_neutral_loss(void* h, const char* group, float mass) {
    ((RefScorer*)h)->peptide.addNeutralLoss(group, mass);
}

// Ascore.pyx:129-152
void refshim_score(void* h, const double* mz, const double* inten, size_t n_peaks,
                   const char* pep, size_t n_of_mod, size_t max_charge,
                   const unsigned int* aux_pos, const float* aux_mass, size_t n_aux) {
    RefScorer* s = (RefScorer*)h;
    s->spectra.consumeSpectra(mz, inten, n_peaks);
    if (aux_pos != nullptr && aux_mass != nullptr)
        s->peptide.consumePeptide(pep, n_of_mod, max_charge, aux_pos, aux_mass, n_aux);
    else
        s->peptide.consumePeptide(pep, n_of_mod, max_charge);
    while (s->spectra.getBin() < s->spectra.getNBins()) {
        s->spectra.resetRank();
        while (s->spectra.getRank() < s->spectra.getNPeaks()) {
            s->peptide.consumePeak(s->spectra.getMZ(), s->spectra.getRank());
            s->spectra.nextRank();
        }
        s->spectra.nextBin();
    }
    s->ascore.score(s->spectra, s->peptide);
}

int refshim_best_sequence(void* h, char* buf, int cap) {
    std::string q = ((RefScorer*)h)->ascore.getBestSequence();
    int n = (int)q.size();
    if (n + 1 > cap) return -n - 1;
    memcpy(buf, q.c_str(), n + 1);
    return n;
}

float refshim_best_score(void* h) { return ((RefScorer*)h)->ascore.getBestScore(); }

size_t refshim_n_pep_scores(void* h) { return ((RefScorer*)h)->ascore.getAllPepScores().size(); }

size_t refshim_sig_len(void* h) {
    std::vector<ScoreContainer> v = ((RefScorer*)h)->ascore.getAllPepScores();
    return v.empty() ? 0 : v[0].signature.size();
}

// Row-major tables in the reference's pep_scores order.
void refshim_pep_scores(void* h, int32_t* signature, int32_t* counts, float* scores,
                        float* weighted, int64_t* total_fragments) {
    std::vector<ScoreContainer> v = ((RefScorer*)h)->ascore.getAllPepScores();
    for (size_t i = 0; i < v.size(); i++) {
        size_t S = v[i].signature.size(), D = v[i].counts.size();
        for (size_t j = 0; j < S; j++) signature[i * S + j] = (int32_t)v[i].signature[j];
        for (size_t j = 0; j < D; j++) {
            counts[i * D + j] = (int32_t)v[i].counts[j];
            scores[i * D + j] = v[i].scores[j];
        }
        weighted[i] = v[i].weighted_score;
        total_fragments[i] = (int64_t)v[i].total_fragments;
    }
}

// '\n'-joined sequences in pep_scores order; returns bytes needed (incl. NUL).
long refshim_sequences(void* h, char* buf, long cap) {
    std::vector<std::string> v = ((RefScorer*)h)->ascore.getAllSequences();
    std::string all;
    for (size_t i = 0; i < v.size(); i++) { if (i) all += '\n'; all += v[i]; }
    long need = (long)all.size() + 1;
    if (need <= cap) memcpy(buf, all.c_str(), need);
    return need;
}

size_t refshim_n_ascores(void* h) { return ((RefScorer*)h)->ascore.getAscores().size(); }

void refshim_ascores(void* h, float* out) {
    std::vector<float> v = ((RefScorer*)h)->ascore.getAscores();
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
}

// Ascore.pyx:266-288 (loop bound is getNumberOfMods there; the caller does the loop)
long refshim_alt_sites(void* h, size_t site, uint32_t* out, long cap) {
    std::vector<size_t> v = ((RefScorer*)h)->ascore.getAlternativeSites(site);
    long n = (long)v.size();
    for (long i = 0; i < n && i < cap; i++) out[i] = (uint32_t)v[i];
    return n;
}

float refshim_calculate_ambiguity(void* h, size_t S, size_t D,
                                  const int32_t* sig_a, const int32_t* cnt_a, const float* sc_a,
                                  float w_a, int64_t tot_a,
                                  const int32_t* sig_b, const int32_t* cnt_b, const float* sc_b,
                                  float w_b, int64_t tot_b) {
    ScoreContainer a, b;
    for (size_t i = 0; i < S; i++) { a.signature.push_back(sig_a[i]); b.signature.push_back(sig_b[i]); }
    for (size_t i = 0; i < D; i++) {
        a.counts.push_back(cnt_a[i]); a.scores.push_back(sc_a[i]);
        b.counts.push_back(cnt_b[i]); b.scores.push_back(sc_b[i]);
    }
    a.weighted_score = w_a; a.total_fragments = tot_a;
    b.weighted_score = w_b; b.total_fragments = tot_b;
    return ((RefScorer*)h)->ascore.calculateAmbiguity(a, b);
}

// ---- stage probes -------------------------------------------------------

// Retained peaks of the last consumed spectrum: (bin, rank, mz, intensity); returns count.
long refshim_binned(void* h, int32_t* bin, int32_t* rank, double* mz, double* inten, long cap,
                    float* min_mz, float* max_mz, int64_t* n_bins) {
    RefScorer* s = (RefScorer*)h;
    *min_mz = s->spectra.getMinMZ(); *max_mz = s->spectra.getMaxMZ();
    *n_bins = (int64_t)s->spectra.getNBins();
    long n = 0;
    s->spectra.resetBin();
    while (s->spectra.getBin() < s->spectra.getNBins()) {
        s->spectra.resetRank();
        while (s->spectra.getRank() < s->spectra.getNPeaks()) {
            if (n < cap) {
                bin[n] = (int32_t)s->spectra.getBin(); rank[n] = (int32_t)s->spectra.getRank();
                mz[n] = s->spectra.getMZ(); inten[n] = s->spectra.getIntensity();
            }
            n++;
            s->spectra.nextRank();
        }
        s->spectra.nextBin();
    }
    s->spectra.resetBin(); s->spectra.resetRank();
    return n;
}

void refshim_consume_spectra(void* h, const double* mz, const double* inten, size_t n) {
    ((RefScorer*)h)->spectra.consumeSpectra(mz, inten, n);
}

void refshim_consume_peptide(void* h, const char* pep, size_t n_of_mod, size_t max_charge,
                             const unsigned int* aux_pos, const float* aux_mass, size_t n_aux) {
    RefScorer* s = (RefScorer*)h;
    if (aux_pos != nullptr && aux_mass != nullptr)
        s->peptide.consumePeptide(pep, n_of_mod, max_charge, aux_pos, aux_mass, n_aux);
    else
        s->peptide.consumePeptide(pep, n_of_mod, max_charge);
}

// Walk the whole fragment graph of (type, charge): for each signature in the
// reference's enumeration order emit its signature bits (N->C) and every fragment m/z.
// sig_out: n_sig x S, frag_off: n_sig+1, frag_out: all fragments. Returns n_sig, or -needed.
long refshim_fragment_graph(void* h, char type, size_t charge, int32_t* sig_out, long sig_cap,
                            int64_t* frag_off, float* frag_out, long frag_cap, int64_t* n_frag_total) {
    RefScorer* s = (RefScorer*)h;
    long n_sig = 0; int64_t nf = 0;
    for (ModifiedPeptide::FragmentGraph g = s->peptide.getFragmentGraph(type, charge);
         !g.isSignatureEnd(); g.incrSignature()) {
        std::vector<size_t> sig = g.getSignature();
        // incrSignature() resumes mid-peptide; restart to list every fragment of this isoform.
        ModifiedPeptide::FragmentGraph g2 = s->peptide.getFragmentGraph(type, charge);
        g2.setSignature(sig);
        if (n_sig < sig_cap) {
            for (size_t j = 0; j < sig.size(); j++) sig_out[n_sig * sig.size() + j] = (int32_t)sig[j];
            frag_off[n_sig] = nf;
        }
        for (; !g2.isFragmentEnd(); g2.incrFragment()) {
            if (nf < frag_cap) frag_out[nf] = g2.getFragmentMZ();
            nf++;
        }
        n_sig++;
    }
    if (n_sig <= sig_cap) frag_off[n_sig < sig_cap ? n_sig : sig_cap] = nf;
    *n_frag_total = nf;
    return n_sig;
}

long refshim_site_determining(void* h, size_t S, const int32_t* sig_a, const int32_t* sig_b,
                              char type, size_t max_charge, float* out_a, long* n_a,
                              float* out_b, long* n_b, long cap) {
    std::vector<size_t> a(sig_a, sig_a + S), b(sig_b, sig_b + S);
    std::vector<std::vector<float>> ions =
        ((RefScorer*)h)->peptide.getSiteDeterminingIons(a, b, type, max_charge);
    *n_a = (long)ions[0].size(); *n_b = (long)ions[1].size();
    for (long i = 0; i < *n_a && i < cap; i++) out_a[i] = ions[0][i];
    for (long i = 0; i < *n_b && i < cap; i++) out_b[i] = ions[1][i];
    return 0;
}

int refshim_get_peptide(void* h, size_t S, const int32_t* sig, char* buf, int cap) {
    std::vector<size_t> v(sig, sig + S);
    std::string q = ((RefScorer*)h)->peptide.getPeptide(v);
    int n = (int)q.size();
    if (n + 1 > cap) return -n - 1;
    memcpy(buf, q.c_str(), n + 1);
    return n;
}

int refshim_has_match(void* h, float mz, float* peak_mz, int64_t* rank) {
    RefScorer* s = (RefScorer*)h;
    if (!s->peptide.hasMatch(mz)) return 0;
    std::tuple<float, size_t> m = s->peptide.getMatch(mz);
    *peak_mz = std::get<0>(m); *rank = (int64_t)std::get<1>(m);
    return 1;
}

// ---- math helpers (cpp/Util.cpp) ------------------------------------------
float refshim_log_sum(float a, float b) { LogMath m; return m.log_sum(a, b); }
float refshim_log_bin_coef(size_t k, size_t n) { LogMath m; return m.log_bin_coef(k, n); }
void* refshim_binom_new(float p) { return new BinomialDist(p); }
void refshim_binom_free(void* b) { delete (BinomialDist*)b; }
float refshim_binom_log_pmf(void* b, size_t k, size_t n) { return ((BinomialDist*)b)->log_pmf(k, n); }
float refshim_binom_log_pvalue(void* b, size_t k, size_t n) { return ((BinomialDist*)b)->log_pvalue(k, n); }
float refshim_binom_log10_pvalue(void* b, size_t k, size_t n) { return ((BinomialDist*)b)->log10_pvalue(k, n); }
long refshim_power_set_sum(const float* v, size_t n, size_t depth, float* out, long cap) {
    std::vector<float> t(v, v + n);
    PowerSetSum p(t, depth);
    long c = 0;
    for (;;) {
        if (c < cap) out[c] = p.getSum();
        c++;
        if (!p.hasNext()) break;
        p.next();
    }
    return c;
}

// ---- batched scoring over CSR arrays (CPU baseline timing; one object, serial) ------
// Same per-PSM sequence as refshim_score; result summaries only.
void refshim_score_batch(void* h, int64_t n_psm, const int64_t* spec_off, const double* mz,
                         const double* inten, const int32_t* psm_spec, const int32_t* pep_off,
                         const char* pep, const int32_t* n_mod, const int32_t* max_charge,
                         const int32_t* aux_off, const uint32_t* aux_pos, const float* aux_mass,
                         float* best_score, float* ascores /* n_psm x max_k */, int32_t max_k) {
    RefScorer* s = (RefScorer*)h;
    std::string p;
    for (int64_t i = 0; i < n_psm; i++) {
        int64_t sp = psm_spec[i];
        p.assign(pep + pep_off[i], pep + pep_off[i + 1]);
        int32_t na = aux_off ? aux_off[i + 1] - aux_off[i] : 0;
        refshim_score(h, mz + spec_off[sp], inten + spec_off[sp], (size_t)(spec_off[sp + 1] - spec_off[sp]),
                      p.c_str(), (size_t)n_mod[i], (size_t)max_charge[i],
                      na ? aux_pos + aux_off[i] : nullptr, na ? aux_mass + aux_off[i] : nullptr, (size_t)na);
        best_score[i] = s->ascore.getBestScore();
        if (ascores) {
            std::vector<float> a = s->ascore.getAscores();
            for (int32_t j = 0; j < max_k; j++)
                ascores[i * max_k + j] = j < (int32_t)a.size() ? a[j] : 0.f;
        }
    }
}

}  // extern "C"
