// setup.cpp — host-side mirror of the set-up the reference performs BEFORE the hot path, so the
// engine can run stand-alone (bench, tests) and so a maintainer can check it against the
// reference's own values.  In a true drop-in the reference host computes all of this and hands
// the arrays to bn_query_load directly.
//
// Mirrors (reference file:line; core/ = c++/src/algo/blast/core):
//   query concatenation / contexts      inc-core/blast_query_info.h:46-58, SURVEY.md A.2
//   mask -> lookup segments             BLAST_ComplementMaskLocations core/blast_filter.c:1019-1119,
//                                       BlastSeqLocCombine :974-1016, BlastSetUp_MaskQuery :1343
//   score matrix / table                BlastScoreBlkNuclMatrixCreate core/blast_stat.c:1036-1105,
//                                       core/blast_parameters.c:236-261
//   ungapped Karlin-Altschul block      Blast_ScoreBlkKbpUngappedCalc core/blast_stat.c:2711-2807,
//                                       Blast_KarlinBlkUngappedCalc :2673, Blast_KarlinLambdaNR :2541,
//                                       BlastKarlinLtoH :2581, BlastKarlinLHtoK :2221
//   gapped Karlin-Altschul block        Blast_KarlinBlkNuclGappedCalc core/blast_stat.c:3806-3905 and
//                                       the published parameter tables :595-704
//   effective lengths                   BLAST_CalcEffLengths core/blast_setup.c:635-790,
//                                       BLAST_ComputeLengthAdjustment core/blast_stat.c:4994-5080
//   cutoffs / X-drops                   core/blast_parameters.c:161-418, :420-480, :760-975
//   lookup-table choice and fill        BlastChooseNaLookupTable core/blast_nalookup.c:51-189,
//                                       BlastMBLookupTableNew :941-1027, s_FillContigMBTable :832-937,
//                                       BlastSmallNaLookupTableNew :384-425, s_BlastSmallNaLookupFinalize
//                                       :200-306, BlastLookupIndexQueryExactMatches core/blast_lookup.c:84-138
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include "../../include/gblastn_b200.h"

namespace {

// Zero-filled host array; page-locked when a CUDA driver is present so bn_query_load's H2D copies
// run at full PCIe speed, plain calloc otherwise (CPU-only set-up tests).
template <typename T>
struct HostArray {
    T *p = nullptr;
    size_t n = 0;
    bool pinned = false;
    HostArray() = default;
    HostArray(const HostArray &) = delete;
    HostArray &operator=(const HostArray &) = delete;
    ~HostArray() { release(); }
    void release()
    {
        if (p) { if (pinned) cudaFreeHost(p); else free(p); }
        p = nullptr; n = 0; pinned = false;
    }
    void assign_zero(size_t count)
    {
        release();
        if (count == 0) return;
        void *q = nullptr;
        if (count * sizeof(T) >= (1u << 20) && cudaHostAlloc(&q, count * sizeof(T), cudaHostAllocDefault) == cudaSuccess) {
            pinned = true;
            memset(q, 0, count * sizeof(T));
        } else {
            (void)cudaGetLastError();
            q = calloc(count, sizeof(T));
        }
        p = (T *)q; n = count;
    }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
    T *data() { return p; }
    const T *data() const { return p; }
    bool empty() const { return n == 0; }
    size_t size() const { return n; }
};

const double kLn2 = 0.69314718055994530941723212145818;

struct KBlk { double Lambda = -1, K = -1, logK = 0, H = -1; };

struct Range { int32_t left, right; };

// ------------------------------------------------------------------ Karlin-Altschul, ungapped
int32_t gcd(int32_t a, int32_t b)
{
    b = std::abs(b);
    if (b > a) std::swap(a, b);
    while (b != 0) { int32_t c = a % b; a = b; b = c; }
    return a;
}

double powi(double x, int32_t n)
{
    if (n == 0) return 1.0;
    if (x == 0.0) return n < 0 ? HUGE_VAL : 0.0;
    if (n < 0) { x = 1.0 / x; n = -n; }
    double y = 1.0;
    while (n > 0) { if (n & 1) y *= x; n /= 2; x *= x; }
    return y;
}

double expm1_series(double x)
{
    const double a = std::fabs(x);
    if (a > .33) return std::exp(x) - 1.;
    if (a < 1.e-16) return x;
    return x * (1. + x * (1. / 2. + x * (1. / 6. + x * (1. / 24. + x * (1. / 120. + x * (1. / 720. +
           x * (1. / 5040. + x * (1. / 40320. + x * (1. / 362880. + x * (1. / 3628800. +
           x * (1. / 39916800. + x * (1. / 479001600. + x / 6227020800.))))))))))));
}

// score-frequency table indexed by score - lo
struct ScoreFreq {
    int32_t lo = 0, hi = 0;            // allocated range
    int32_t obs_min = 0, obs_max = 0;
    double score_avg = 0;
    std::vector<double> p;
    double &at(int32_t s) { return p[(size_t)(s - lo)]; }
    double at(int32_t s) const { return p[(size_t)(s - lo)]; }
};

double lambda_nr(const ScoreFreq &sf)
{
    const int32_t low = sf.obs_min, high = sf.obs_max;
    if (sf.score_avg >= 0.) return -1.0;
    if (low >= 0 || high <= 0) return -1.0;
    int32_t d = -low;
    for (int32_t i = 1; i <= high - low && d > 1; ++i)
        if (sf.at(i + low) != 0.0) d = gcd(d, i);
    const double lambda0 = 0.5, tolx = 1.e-5;
    const int32_t itmax = 20, maxNewton = 20 + 17;
    double x0 = std::exp(-lambda0);
    double x = (0 < x0 && x0 < 1) ? x0 : .5;
    double a = 0, b = 1, f = 4;
    bool isNewton = false;
    for (int32_t k = 0; k < itmax; k++) {
        double g = 0, fold = f;
        const bool wasNewton = isNewton;
        isNewton = false;
        f = sf.at(low);
        int32_t i;
        for (i = low + d; i < 0; i += d) { g = x * g + f; f = f * x + sf.at(i); }
        g = x * g + f;
        f = f * x + sf.at(0) - 1;
        for (i = d; i <= high; i += d) { g = x * g + f; f = f * x + sf.at(i); }
        if (f > 0) a = x;
        else if (f < 0) b = x;
        else break;
        if (b - a < 2 * a * (1 - b) * tolx) { x = (a + b) / 2; break; }
        if (k >= maxNewton || (wasNewton && std::fabs(f) > .9 * std::fabs(fold)) || g >= 0) {
            x = (a + b) / 2;
        } else {
            const double p = -f / g, y = x + p;
            if (y <= a || y >= b) x = (a + b) / 2;
            else {
                isNewton = true;
                x = y;
                if (std::fabs(p) < tolx * x * (1 - x)) break;
            }
        }
    }
    return -std::log(x) / d;
}

double l_to_h(const ScoreFreq &sf, double lambda)
{
    if (lambda < 0.) return -1.;
    const int32_t low = sf.obs_min, high = sf.obs_max;
    if (low >= 0 || high <= 0) return -1.;
    const double etonlam = std::exp(-lambda);
    double sum = low * sf.at(low);
    for (int32_t s = low + 1; s <= high; s++) sum = s * sf.at(s) + etonlam * sum;
    const double scale = powi(etonlam, high);
    if (scale > 0.0) return lambda * sum / scale;
    return lambda * std::exp(lambda * high + std::log(sum));
}

double lh_to_k(const ScoreFreq &sf, double lambda, double H)
{
    if (lambda <= 0. || H <= 0.) return -1.;
    if (sf.score_avg >= 0.0) return -1.;
    int32_t low = sf.obs_min, high = sf.obs_max;
    int32_t range = high - low;
    int32_t divisor = -low;
    for (int32_t i = 1; i <= range && divisor > 1; ++i)
        if (sf.at(low + i) != 0.0) divisor = gcd(divisor, i);
    high /= divisor; low /= divisor; lambda *= divisor;
    range = high - low;
    double first = H / lambda;
    const double expMinusLambda = std::exp(-lambda);
    if (low == -1 && high == 1) {
        const double pl = sf.at(low * divisor), ph = sf.at(high * divisor);
        return (pl - ph) * (pl - ph) / pl;
    }
    if (low == -1 || high == 1) {
        if (high != 1) {
            const double avg = sf.score_avg / divisor;
            first = (avg * avg) / first;
        }
        return first * (1.0 - expMinusLambda);
    }
    const double sumlimit = 0.0001;
    const int iterlimit = 100;
    std::vector<double> P((size_t)(iterlimit * range + 1), 0.0);
    // probabilities re-indexed so that (low*divisor) is at index 0, stride `divisor`... the
    // reference indexes probArrayStartLow[i] with i in units of the ORIGINAL score spacing only
    // when divisor == 1; for divisor > 1 it reads consecutive entries, as restated here.
    auto prob = [&](int32_t i) { return sf.at(sf.obs_min + i); };
    double outerSum = 0., innerSum = 1., oldsum = 1., oldsum2 = 1.;
    int32_t lowAl = 0, highAl = 0;
    P[0] = 1.;
    int iter = 0;
    while (iter < iterlimit && innerSum > sumlimit) {
        int32_t first_i = range, last_i = range;
        lowAl += low; highAl += high;
        for (int32_t pp = highAl - lowAl; pp >= 0; --pp) {
            int32_t p1 = pp - first_i, p1e = pp - last_i, p2 = first_i;
            innerSum = 0.;
            while (p1 >= p1e) { innerSum += P[(size_t)p1] * prob(p2); --p1; ++p2; }
            if (first_i) --first_i;
            if (pp <= range) --last_i;
            P[(size_t)pp] = innerSum;
        }
        int32_t idx = 0;
        innerSum = P[0];
        int32_t i;
        for (i = lowAl + 1; i < 0; i++) innerSum = P[(size_t)++idx] + innerSum * expMinusLambda;
        innerSum *= expMinusLambda;
        for (; i <= highAl; ++i) innerSum += P[(size_t)++idx];
        oldsum2 = oldsum; oldsum = innerSum;
        ++iter;
        innerSum /= iter;
        outerSum += innerSum;
    }
    (void)oldsum2;
    return -std::exp(-2.0 * outerSum) / (first * expm1_series(-lambda));
}

bool ungapped_kbp(const ScoreFreq &sf, KBlk &k)
{
    k.Lambda = lambda_nr(sf);
    if (k.Lambda < 0.) return false;
    k.H = l_to_h(sf, k.Lambda);
    if (k.H < 0.) return false;
    k.K = lh_to_k(sf, k.Lambda, k.H);
    if (k.K < 0.) return false;
    k.logK = std::log(k.K);
    return true;
}

// ------------------------------------------------------------------ gapped parameter tables
struct Row8 { double v[8]; };   // gap_open, gap_extend, Lambda, K, H, alpha, beta, theta
struct NuclTable { int reward, penalty; std::vector<Row8> rows; int open_max, extend_max; bool round_down; };

const std::vector<NuclTable> &nucl_tables()
{
    static const std::vector<NuclTable> t = {
        {1, -5, {{{0, 0, 1.39, 0.747, 1.38, 1.00, 0, 100}}, {{3, 3, 1.39, 0.747, 1.38, 1.00, 0, 100}}}, 3, 3, false},
        {1, -4, {{{0, 0, 1.383, 0.738, 1.36, 1.02, 0, 100}}, {{1, 2, 1.36, 0.67, 1.2, 1.1, 0, 98}},
                 {{0, 2, 1.26, 0.43, 0.90, 1.4, -1, 91}}, {{2, 1, 1.35, 0.61, 1.1, 1.2, -1, 98}},
                 {{1, 1, 1.22, 0.35, 0.72, 1.7, -3, 88}}}, 2, 2, false},
        {2, -7, {{{0, 0, 0.69, 0.73, 1.34, 0.515, 0, 100}}, {{2, 4, 0.68, 0.67, 1.2, 0.55, 0, 99}},
                 {{0, 4, 0.63, 0.43, 0.90, 0.7, -1, 91}}, {{4, 2, 0.675, 0.62, 1.1, 0.6, -1, 98}},
                 {{2, 2, 0.61, 0.35, 0.72, 1.7, -3, 88}}}, 4, 4, true},
        {1, -3, {{{0, 0, 1.374, 0.711, 1.31, 1.05, 0, 100}}, {{2, 2, 1.37, 0.70, 1.2, 1.1, 0, 99}},
                 {{1, 2, 1.35, 0.64, 1.1, 1.2, -1, 98}}, {{0, 2, 1.25, 0.42, 0.83, 1.5, -2, 91}},
                 {{2, 1, 1.34, 0.60, 1.1, 1.2, -1, 97}}, {{1, 1, 1.21, 0.34, 0.71, 1.7, -2, 88}}}, 2, 2, false},
        {2, -5, {{{0, 0, 0.675, 0.65, 1.1, 0.6, -1, 99}}, {{2, 4, 0.67, 0.59, 1.1, 0.6, -1, 98}},
                 {{0, 4, 0.62, 0.39, 0.78, 0.8, -2, 91}}, {{4, 2, 0.67, 0.61, 1.0, 0.65, -2, 98}},
                 {{2, 2, 0.56, 0.32, 0.59, 0.95, -4, 82}}}, 4, 4, true},
        {1, -2, {{{0, 0, 1.28, 0.46, 0.85, 1.5, -2, 96}}, {{2, 2, 1.33, 0.62, 1.1, 1.2, 0, 99}},
                 {{1, 2, 1.30, 0.52, 0.93, 1.4, -2, 97}}, {{0, 2, 1.19, 0.34, 0.66, 1.8, -3, 89}},
                 {{3, 1, 1.32, 0.57, 1.0, 1.3, -1, 99}}, {{2, 1, 1.29, 0.49, 0.92, 1.4, -1, 96}},
                 {{1, 1, 1.14, 0.26, 0.52, 2.2, -5, 85}}}, 2, 2, false},
        {2, -3, {{{0, 0, 0.55, 0.21, 0.46, 1.2, -5, 87}}, {{4, 4, 0.63, 0.42, 0.84, 0.75, -2, 99}},
                 {{2, 4, 0.615, 0.37, 0.72, 0.85, -3, 97}}, {{0, 4, 0.55, 0.21, 0.46, 1.2, -5, 87}},
                 {{3, 3, 0.615, 0.37, 0.68, 0.9, -3, 97}}, {{6, 2, 0.63, 0.42, 0.84, 0.75, -2, 99}},
                 {{5, 2, 0.625, 0.41, 0.78, 0.8, -2, 99}}, {{4, 2, 0.61, 0.35, 0.68, 0.9, -3, 96}},
                 {{2, 2, 0.515, 0.14, 0.33, 1.55, -9, 81}}}, 6, 4, true},
        {3, -4, {{{6, 3, 0.389, 0.25, 0.56, 0.7, -5, 95}}, {{5, 3, 0.375, 0.21, 0.47, 0.8, -6, 92}},
                 {{4, 3, 0.351, 0.14, 0.35, 1.0, -9, 86}}, {{6, 2, 0.362, 0.16, 0.45, 0.8, -4, 88}},
                 {{5, 2, 0.330, 0.092, 0.28, 1.2, -13, 81}}, {{4, 2, 0.281, 0.046, 0.16, 1.8, -23, 69}}}, 6, 3, true},
        {4, -5, {{{0, 0, 0.22, 0.061, 0.22, 1.0, -15, 74}}, {{6, 5, 0.28, 0.21, 0.47, 0.6, -7, 93}},
                 {{5, 5, 0.27, 0.17, 0.39, 0.7, -9, 90}}, {{4, 5, 0.25, 0.10, 0.31, 0.8, -10, 83}},
                 {{3, 5, 0.23, 0.065, 0.25, 0.9, -11, 76}}}, 12, 8, false},
        {1, -1, {{{3, 2, 1.09, 0.31, 0.55, 2.0, -2, 99}}, {{2, 2, 1.07, 0.27, 0.49, 2.2, -3, 97}},
                 {{1, 2, 1.02, 0.21, 0.36, 2.8, -6, 92}}, {{0, 2, 0.80, 0.064, 0.17, 4.8, -16, 72}},
                 {{4, 1, 1.08, 0.28, 0.54, 2.0, -2, 98}}, {{3, 1, 1.06, 0.25, 0.46, 2.3, -4, 96}},
                 {{2, 1, 0.99, 0.17, 0.30, 3.3, -10, 90}}}, 4, 2, false},
        {3, -2, {{{5, 5, 0.208, 0.030, 0.072, 2.9, -47, 77}}}, 5, 5, false},
        {5, -4, {{{10, 6, 0.163, 0.068, 0.16, 1.0, -19, 85}}, {{8, 6, 0.146, 0.039, 0.11, 1.3, -29, 76}}}, 25, 10, false},
    };
    return t;
}

struct NuclValues {
    std::vector<Row8> normal;
    bool has_linear = false;
    Row8 linear{};
    int open_max = 0, extend_max = 0;
    bool round_down = false;
};

// s_GetNuclValuesArray core/blast_stat.c:3207-3345 (+ s_SplitArrayOf8, s_AdjustGapParametersByGcd)
bool nucl_values(int reward, int penalty, NuclValues &out)
{
    const int divisor = gcd(reward, penalty);
    if (divisor != 1) { reward /= divisor; penalty /= divisor; }
    for (const NuclTable &t : nucl_tables()) {
        if (t.reward != reward || t.penalty != penalty) continue;
        out.round_down = t.round_down;
        out.open_max = t.open_max; out.extend_max = t.extend_max;
        size_t first = 0;
        if (t.rows[0].v[0] == 0 && t.rows[0].v[1] == 0) { out.has_linear = true; out.linear = t.rows[0]; first = 1; }
        out.normal.assign(t.rows.begin() + (long)first, t.rows.end());
        if (divisor != 1) {
            out.open_max *= divisor; out.extend_max *= divisor;
            for (Row8 &r : out.normal) { r.v[0] *= divisor; r.v[1] *= divisor; r.v[2] /= divisor; r.v[5] /= divisor; }
            if (out.has_linear) { Row8 &r = out.linear; r.v[0] *= divisor; r.v[1] *= divisor; r.v[2] /= divisor; r.v[5] /= divisor; }
        }
        return true;
    }
    return false;
}

// Blast_KarlinBlkNuclGappedCalc
bool gapped_kbp(int gap_open, int gap_extend, int reward, int penalty, const KBlk &ungapped, KBlk &k,
                bool &round_down, std::string &err)
{
    NuclValues nv;
    if (!nucl_values(reward, penalty, nv)) { err = "substitution scores are not supported"; return false; }
    round_down = nv.round_down;
    if (gap_open == 0 && gap_extend == 0 && nv.has_linear) {
        k.Lambda = nv.linear.v[2]; k.K = nv.linear.v[3]; k.logK = std::log(k.K); k.H = nv.linear.v[4];
        return true;
    }
    for (const Row8 &r : nv.normal)
        if (r.v[0] == gap_open && r.v[1] == gap_extend) {
            k.Lambda = r.v[2]; k.K = r.v[3]; k.logK = std::log(k.K); k.H = r.v[4];
            return true;
        }
    if (gap_open >= nv.open_max && gap_extend >= nv.extend_max) { k = ungapped; return true; }
    err = "gap existence / extension values are not supported for these substitution scores";
    return false;
}

// Blast_GetNuclAlphaBeta (gapped search)
void nucl_alpha_beta(int reward, int penalty, int gap_open, int gap_extend, const KBlk &ungapped,
                     double &alpha, double &beta)
{
    NuclValues nv;
    bool found = false;
    if (nucl_values(reward, penalty, nv) && !nv.normal.empty()) {
        if (gap_open == 0 && gap_extend == 0 && nv.has_linear) { alpha = nv.linear.v[5]; beta = nv.linear.v[6]; found = true; }
        else for (const Row8 &r : nv.normal)
            if (r.v[0] == gap_open && r.v[1] == gap_extend) { alpha = r.v[5]; beta = r.v[6]; found = true; break; }
    }
    if (!found) {
        alpha = ungapped.Lambda / ungapped.H;
        beta = ((reward == 1 && penalty == -1) || (reward == 2 && penalty == -3)) ? -2 : 0;
    }
}

// BLAST_ComputeLengthAdjustment
int32_t length_adjustment(double K, double logK, double alpha_d_lambda, double beta, int32_t query_length,
                          int64_t db_length, int32_t db_num_seqs)
{
    const double m = (double)query_length, n = (double)db_length, N = (double)db_num_seqs;
    double ell, ss, ell_min = 0, ell_max, ell_next = 0;
    bool converged = false;
    {
        const double a = N, mb = m * N + n, c = n * m - std::max(m, n) / K;
        if (c < 0) return 0;
        ell_max = 2 * c / (mb + std::sqrt(mb * mb - 4 * a * c));
    }
    for (int i = 1; i <= 20; i++) {
        ell = ell_next;
        ss = (m - ell) * (n - N * ell);
        const double ell_bar = alpha_d_lambda * (logK + std::log(ss)) + beta;
        if (ell_bar >= ell) {
            ell_min = ell;
            if (ell_bar - ell_min <= 1.0) { converged = true; break; }
            if (ell_min == ell_max) break;
        } else ell_max = ell;
        if (ell_min <= ell_bar && ell_bar <= ell_max) ell_next = ell_bar;
        else ell_next = (i == 1) ? ell_max : (ell_min + ell_max) / 2;
    }
    int32_t result = (int32_t)ell_min;
    if (converged) {
        ell = std::ceil(ell_min);
        if (ell <= ell_max) {
            ss = (m - ell) * (n - N * ell);
            if (alpha_d_lambda * (logK + std::log(ss)) + beta >= ell) result = (int32_t)ell;
        }
    }
    return result;
}

// BlastKarlinEtoS_simple
int32_t e_to_s(double E, const KBlk &k, int64_t searchsp)
{
    if (k.Lambda < 0. || k.K < 0. || k.H < 0.0) return SHRT_MIN;
    E = std::max(E, 1.0e-297);
    return (int32_t)std::ceil(std::log((double)(k.K * searchsp / E)) / k.Lambda);
}

}  // namespace

// ================================================================================================
struct BnSetup {
    BnQueryBatch batch{};
    HostArray<uint8_t> query;
    std::vector<BnContext> ctx;
    HostArray<int32_t> hashtable, next_pos;
    std::vector<int32_t> masked, segments;
    std::vector<uint32_t> pv;
    std::vector<int16_t> backbone, overflow;
    std::vector<int32_t> na_backbone, na_overflow;
    std::vector<double> kbp_std, kbp_gap;
    int32_t gap_x_dropoff_final = 0, longest_chain = 0;
    std::string error;
};

namespace {

const uint8_t kComplement[16] = {3, 2, 1, 0, 5, 4, 7, 6, 8, 9, 13, 12, 11, 10, 14, 15};
const uint8_t kBlastnaToNcbi4na[16] = {1, 2, 4, 8, 5, 10, 3, 12, 9, 6, 14, 13, 11, 7, 15, 0};

// BlastSeqLocCombine(link_value = 0): sort by start, merge overlapping
void combine(std::vector<Range> &v)
{
    if (v.empty()) return;
    std::stable_sort(v.begin(), v.end(), [](const Range &a, const Range &b) {
        if (a.left != b.left) return a.left < b.left;
        return a.right < b.right;
    });
    std::vector<Range> out;
    out.push_back(v[0]);
    for (size_t i = 1; i < v.size(); i++) {
        Range &t = out.back();
        if (t.right > v[i].left) t.right = std::max(t.right, v[i].right);
        else out.push_back(v[i]);
    }
    v.swap(out);
}

// BLAST_ComplementMaskLocations for one context; `mask` in plus-strand query coordinates.
void complement_context(const BnContext &c, std::vector<Range> mask, bool reverse, std::vector<Range> &segs)
{
    const int32_t start_offset = c.query_offset, end_offset = c.query_offset + c.query_length - 1;
    if (mask.empty()) { segs.push_back(Range{start_offset, end_offset}); return; }
    if (reverse) std::reverse(mask.begin(), mask.end());
    bool first = true, open = true;
    int32_t left = 0, right;
    for (const Range &m : mask) {
        int32_t fs, fe;
        if (reverse) { fs = end_offset - m.right; fe = end_offset - m.left; }
        else { fs = start_offset + m.left; fe = start_offset + m.right; }
        if (first) {
            open = true; first = false;
            if (fs > start_offset) left = start_offset;
            else { left = fe + 1; continue; }
        }
        right = fs - 1;
        segs.push_back(Range{left, right});
        if (fe >= end_offset) { open = false; break; }
        left = fe + 1;
    }
    if (open) segs.push_back(Range{left, end_offset});
}

// BlastChooseNaLookupTable (with G-BLASTN's word-size-11 edit, core/blast_nalookup.c:127-144).
// returns lut type (0 MB, 1 SmallNa, 2 Na)
int choose_table(int word_size, int32_t entries, int32_t max_q_off, int &lut_width)
{
    int type;
    switch (word_size) {
    case 4: case 5: case 6: type = 1; lut_width = word_size; break;
    case 7: type = 1; lut_width = entries < 250 ? 6 : 7; break;
    case 8: type = 1; lut_width = entries < 8500 ? 7 : 8; break;
    case 9:
        if (entries < 1250) { lut_width = 7; type = 1; }
        else if (entries < 21000) { lut_width = 8; type = 1; }
        else { lut_width = 9; type = 0; }
        break;
    case 10:
        if (entries < 1250) { lut_width = 7; type = 1; }
        else if (entries < 8500) { lut_width = 8; type = 1; }
        else if (entries < 18000) { lut_width = 9; type = 0; }
        else { lut_width = 10; type = 0; }
        break;
    case 11:
        if (entries < 12000) { lut_width = 8; type = 1; }
        else { lut_width = 11; type = 0; }
        break;
    case 12:
        if (entries < 8500) { lut_width = 8; type = 1; }
        else if (entries < 18000) { lut_width = 9; type = 0; }
        else if (entries < 60000) { lut_width = 10; type = 0; }
        else if (entries < 900000) { lut_width = 11; type = 0; }
        else { lut_width = 12; type = 0; }
        break;
    default:
        if (entries < 8500) { lut_width = 8; type = 1; }
        else if (entries < 300000) { lut_width = 11; type = 0; }
        else { lut_width = 12; type = 0; }
        break;
    }
    if (type == 1 && (entries >= 32767 || max_q_off >= 32768)) type = 2;
    return type;
}

int ilog2(int64_t x) { int l = 0; while (x > 1) { x >>= 1; ++l; } return l; }

// s_SeqLocListInvert core/blast_nalookup.c:318-355
void invert_locations(const std::vector<Range> &segs, int32_t length, std::vector<int32_t> &out)
{
    if (segs.empty()) return;
    int32_t start = 0, stop = std::max(0, segs[0].left - 1);
    if (stop - start > 2) { out.push_back(start); out.push_back(stop); }
    for (size_t i = 0; i < segs.size(); i++) {
        start = segs[i].right + 1;
        stop = (i + 1 < segs.size()) ? segs[i + 1].left - 1 : length - 1;
        if (stop - start > 2) { out.push_back(start); out.push_back(stop); }
    }
}

}  // namespace

extern "C" {

int bn_setup_create(const BnSetupOptions *opt, int32_t nq, const uint8_t *qseq, const int32_t *qlens,
                    const int32_t *qmask_n, const int32_t *qmask_iv, BnSetup **out)
{
    if (!opt || !qseq || !qlens || nq <= 0 || !out) return BN_ERR_INVALID;
    *out = nullptr;
    BnSetup *S = new BnSetup();
    BnQueryBatch &b = S->batch;
    const bool mb = opt->task == 0;
    const int word_size = opt->word_size ? opt->word_size : (mb ? 28 : 11);
    const int reward = opt->reward ? opt->reward : (mb ? 1 : 2);
    const int penalty = opt->penalty ? opt->penalty : (mb ? -2 : -3);
    const int gap_open = opt->gap_open >= 0 ? opt->gap_open : (mb ? 0 : 5);
    const int gap_extend = opt->gap_extend >= 0 ? opt->gap_extend : (mb ? 0 : 2);
    const int greedy = opt->greedy >= 0 ? opt->greedy : (mb ? 1 : 0);
    const double xdrop_ungap = opt->xdrop_ungap > 0 ? opt->xdrop_ungap : 20.0;
    const double xdrop_gap = opt->xdrop_gap > 0 ? opt->xdrop_gap : (greedy ? 25.0 : 30.0);
    const double xdrop_gap_final = opt->xdrop_gap_final > 0 ? opt->xdrop_gap_final : 100.0;
    const double evalue = opt->evalue > 0 ? opt->evalue : 10.0;
    const double gap_trigger_bits = 27.0;     // BLAST_GAP_TRIGGER_NUCL
    auto bail = [&](int code) { delete S; return code; };
    if (word_size < 4) return bail(BN_ERR_INVALID);

    // ---- concatenated query + contexts ---------------------------------------------------------
    int64_t total = 1;
    for (int32_t i = 0; i < nq; i++) { if (qlens[i] <= 0) return bail(BN_ERR_INVALID); total += 2 * ((int64_t)qlens[i] + 1); }
    if (total > INT32_MAX - 16) return bail(BN_ERR_OVERFLOW);
    S->query.assign_zero((size_t)total);
    S->ctx.resize((size_t)2 * nq);
    {
        size_t pos = 0;
        const uint8_t *src = qseq;
        S->query[pos++] = 15;
        for (int32_t i = 0; i < nq; i++) {
            const int32_t L = qlens[i];
            for (int c = 0; c < 2; c++) {
                BnContext &x = S->ctx[(size_t)2 * i + c];
                memset(&x, 0, sizeof x);
                x.query_offset = (int32_t)pos - 1; x.query_length = L; x.query_index = i;
                x.frame = c == 0 ? 1 : -1; x.is_valid = 1;
                if (c == 0) for (int32_t k = 0; k < L; k++) S->query[pos++] = src[k] & 15;
                else for (int32_t k = 0; k < L; k++) S->query[pos++] = kComplement[src[L - 1 - k] & 15];
                S->query[pos++] = 15;
            }
            src += L;
        }
    }
    const int32_t concat_len = (int32_t)total - 2;
    uint8_t *Q = S->query.data() + 1;     // query->sequence

    // ---- masks -> lookup segments (and hard masking when !mask_at_hash) -----------------------
    std::vector<Range> segs;
    {
        const int32_t *iv = qmask_iv;
        for (int32_t i = 0; i < nq; i++) {
            std::vector<Range> m;
            if (qmask_n) for (int32_t k = 0; k < qmask_n[i]; k++) m.push_back(Range{iv[2 * k], iv[2 * k + 1]});
            if (qmask_n) iv += 2 * qmask_n[i];
            combine(m);
            for (int c = 0; c < 2; c++) {
                const BnContext &x = S->ctx[(size_t)2 * i + c];
                if (!opt->mask_at_hash)
                    for (const Range &r : m)
                        for (int32_t p = r.left; p <= r.right; p++) {
                            const int32_t pp = c == 0 ? p : x.query_length - 1 - p;
                            if (pp >= 0 && pp < x.query_length) Q[x.query_offset + pp] = 14;   // kNuclMask
                        }
                complement_context(x, m, c == 1, segs);
            }
        }
    }

    // ---- scoring matrix (16 x 16) and the 4-base score table ------------------------------------
    {
        int degeneracy[16];
        for (int i = 0; i < 4; i++) degeneracy[i] = 1;
        for (int i = 4; i < 16; i++) {
            int d = 0;
            for (int j = 0; j < 4; j++) if (kBlastnaToNcbi4na[i] & kBlastnaToNcbi4na[j]) d++;
            degeneracy[i] = d;
        }
        for (int i = 0; i < 16; i++)
            for (int j = i; j < 16; j++) {
                int v;
                if (kBlastnaToNcbi4na[i] & kBlastnaToNcbi4na[j]) {
                    double x = (double)((degeneracy[j] - 1) * penalty + reward) / (double)degeneracy[j];
                    x += (x >= 0. ? 0.5 : -0.5);
                    v = (int)(long)x;
                } else v = penalty;
                b.matrix[16 * i + j] = v; b.matrix[16 * j + i] = v;
            }
        for (int i = 0; i < 16; i++) { b.matrix[16 * 15 + i] = INT_MIN / 2; b.matrix[16 * i + 15] = INT_MIN / 2; }
        for (int i = 0; i < 256; i++) {
            int s = 0;
            s += (i & 3) ? penalty : reward;
            s += ((i >> 2) & 3) ? penalty : reward;
            s += ((i >> 4) & 3) ? penalty : reward;
            s += (i >> 6) ? penalty : reward;
            b.nucl_score_table[i] = s;
        }
    }
    int loscore = SHRT_MAX, hiscore = SHRT_MIN;
    for (int i = 0; i < 256; i++) {
        const int v = b.matrix[i];
        if (v <= SHRT_MIN || v >= SHRT_MAX) continue;
        loscore = std::min(loscore, v); hiscore = std::max(hiscore, v);
    }

    // ---- Karlin-Altschul blocks --------------------------------------------------------------------
    S->kbp_std.assign((size_t)8 * nq, -1.0);
    S->kbp_gap.assign((size_t)8 * nq, -1.0);
    std::vector<KBlk> kstd((size_t)2 * nq), kgap((size_t)2 * nq);
    bool round_down = false;
    for (int32_t c = 0; c < 2 * nq; c++) {
        const BnContext &x = S->ctx[(size_t)c];
        int64_t comp[16] = {0};
        for (int32_t k = 0; k < x.query_length; k++) comp[Q[x.query_offset + k] & 15]++;
        comp[14] = 0; comp[15] = 0;            // BLAST_ScoreSetAmbigRes 'N' and '-'
        double sum = 0, prob1[16], prob2[16] = {0};
        for (int i = 0; i < 16; i++) sum += (double)comp[i];
        for (int i = 0; i < 16; i++) prob1[i] = sum == 0. ? 0.0 : (double)comp[i] / sum;
        for (int i = 0; i < 4; i++) prob2[i] = 25.0 / 100.0;
        ScoreFreq sf;
        sf.lo = loscore; sf.hi = hiscore;
        sf.p.assign((size_t)(hiscore - loscore + 1), 0.0);
        for (int i = 0; i < 16; i++)
            for (int j = 0; j < 16; j++) {
                const int s = b.matrix[16 * i + j];
                if (s >= loscore) sf.at(s) += prob1[i] * prob2[j];
            }
        double score_sum = 0.;
        int obs_min = SHRT_MIN, obs_max = SHRT_MIN;
        for (int s = sf.lo; s <= sf.hi; s++)
            if (sf.at(s) > 0.) { score_sum += sf.at(s); obs_max = s; if (obs_min == SHRT_MIN) obs_min = s; }
        sf.obs_min = obs_min; sf.obs_max = obs_max;
        double avg = 0.0;
        if (score_sum > 0.0001 || score_sum < -0.0001)
            for (int s = obs_min; s <= obs_max; s++) { sf.at(s) /= score_sum; avg += s * sf.at(s); }
        sf.score_avg = avg;
        if (obs_min == SHRT_MIN || !ungapped_kbp(sf, kstd[(size_t)c])) {
            S->error = "could not calculate ungapped Karlin-Altschul parameters for a query";
            return bail(BN_ERR_INVALID);
        }
        std::string err;
        if (!gapped_kbp(gap_open, gap_extend, reward, penalty, kstd[(size_t)c], kgap[(size_t)c], round_down, err))
            return bail(BN_ERR_UNSUPPORTED);
        const KBlk &ks = kstd[(size_t)c], &kg = kgap[(size_t)c];
        double *o = &S->kbp_std[(size_t)4 * c]; o[0] = ks.Lambda; o[1] = ks.K; o[2] = ks.logK; o[3] = ks.H;
        o = &S->kbp_gap[(size_t)4 * c]; o[0] = kg.Lambda; o[1] = kg.K; o[2] = kg.logK; o[3] = kg.H;
    }

    // ---- effective lengths (database search: db_length > 0) -------------------------------------
    for (int32_t c = 0; c < 2 * nq; c++) {
        BnContext &x = S->ctx[(size_t)c];
        double alpha, beta;
        nucl_alpha_beta(reward, penalty, gap_open, gap_extend, kstd[(size_t)c], alpha, beta);
        const KBlk &kg = kgap[(size_t)c];
        const int32_t adj = length_adjustment(kg.K, kg.logK, alpha / kg.Lambda, beta, x.query_length,
                                              opt->db_length, opt->db_num_seqs);
        int64_t eff_db = opt->db_length - (int64_t)opt->db_num_seqs * adj;
        if (eff_db <= 0) eff_db = 1;
        x.length_adjustment = adj;
        x.eff_searchsp = eff_db * (int64_t)(x.query_length - adj);
        x.gap_lambda = kg.Lambda; x.gap_logK = kg.logK;
    }

    // ---- cutoffs -------------------------------------------------------------------------------------
    double min_lambda = (double)INT_MAX;
    for (int32_t c = 0; c < 2 * nq; c++) min_lambda = std::min(min_lambda, kgap[(size_t)c].Lambda);
    b.gap_x_dropoff = (int32_t)(xdrop_gap * kLn2 / min_lambda);
    S->gap_x_dropoff_final = (int32_t)std::max(xdrop_gap_final * kLn2 / min_lambda, (double)b.gap_x_dropoff);
    for (int32_t c = 0; c < 2 * nq; c++) {
        BnContext &x = S->ctx[(size_t)c];
        const KBlk &ks = kstd[(size_t)c];
        int32_t s = 1;
        const int32_t es = e_to_s(evalue, kgap[(size_t)c], x.eff_searchsp);
        if (es > s) s = es;
        x.gapped_cutoff = s;                                    // hit_params cutoff_score (== _max)
        const int32_t x_init = (int32_t)(1.0 * std::ceil(xdrop_ungap * kLn2 / ks.Lambda));
        const int32_t gap_trigger = (int32_t)((gap_trigger_bits * kLn2 + ks.logK) / ks.Lambda);
        const int32_t cutoff = std::min(gap_trigger, x.gapped_cutoff);
        x.cutoff_score = cutoff;
        x.x_dropoff = x_init == 0 ? cutoff : x_init;
        x.reduced_cutoff = (int32_t)(0.9 * cutoff);
    }

    // ---- lookup table ----------------------------------------------------------------------------------
    int32_t entries = 0, max_q_off = 0;
    for (const Range &r : segs) { entries += r.right - r.left; max_q_off = std::max(max_q_off, r.right); }
    int lut_width = 0;
    const int lut_type = choose_table(word_size, entries, max_q_off, lut_width);
    b.word_length = word_size; b.lut_word_length = lut_width;
    b.scan_step = word_size - lut_width + 1;
    const bool mask_at_hash = opt->mask_at_hash != 0;
    if (lut_type == 0) {
        b.lut_type = BN_LUT_MB;
        b.hashsize = (int64_t)1 << (2 * lut_width);
        const bool device_fill = opt->device_lookup != 0;
        if (!device_fill) {
            S->hashtable.assign_zero((size_t)b.hashsize);
            S->next_pos.assign_zero((size_t)concat_len + 1);
        }
        const int64_t kTargetPVSize = 131072;
        int64_t pv_size = b.hashsize <= 8 * kTargetPVSize ? (b.hashsize >> 5) : kTargetPVSize / 4;
        if (entries <= 15000 || entries >= 800000) pv_size /= 2;
        b.pv_array_bts = ilog2(b.hashsize / pv_size);
        S->pv.assign((size_t)pv_size, 0u);
        std::vector<uint32_t> helper((size_t)(b.hashsize / 2048), 0u);
        const int32_t mask = (int32_t)(b.hashsize - 1);
        for (const Range &loc : segs) {
            if (device_fill) break;                 // bn_query_load fills the table on the device
            int32_t from = loc.left;
            const int32_t to = loc.right - lut_width;
            if (word_size > loc.right - loc.left + 1) continue;
            // seq walks sequence_start + from .. ; index = 1-based position of the word start
            const uint8_t *seq = S->query.data() + from;      // == query->sequence_start + from
            const uint8_t *pos = seq + lut_width;
            from -= lut_width - 2;
            const int32_t last_offset = to + 2;
            int32_t ecode = 0;
            for (int32_t index = from; index <= last_offset; index++) {
                const uint8_t val = *++seq;
                if ((val & 0xfc) != 0) { ecode = 0; pos = seq + lut_width; continue; }
                ecode = ((ecode << 2) & mask) + val;
                if (seq < pos) continue;
                if (S->hashtable[(size_t)ecode] == 0)
                    S->pv[(size_t)(ecode >> b.pv_array_bts)] |= 1u << (ecode & 31);
                else helper[(size_t)(ecode / 2048)]++;
                S->next_pos[(size_t)index] = S->hashtable[(size_t)ecode];
                S->hashtable[(size_t)ecode] = index;
            }
        }
        uint32_t longest = 2;
        for (uint32_t h : helper) longest = std::max(longest, h);
        S->longest_chain = (int32_t)longest;
        if (device_fill) S->pv.clear();
    } else {
        // thin backbone of BlastLookupIndexQueryExactMatches (core/blast_lookup.c:84-138): per cell the query
        // offsets of its words in indexing order; shared by the small and the standard blastn table
        b.hashsize = (int64_t)1 << (2 * lut_width);
        std::vector<std::vector<int32_t>> thin((size_t)b.hashsize);
        const int32_t mask = (int32_t)(b.hashsize - 1);
        auto add_word = [&](const uint8_t *w, int32_t q_off) {
            int32_t idx = 0;
            for (int k = 0; k < lut_width; k++) idx = ((idx << 2) | w[k]) & mask;   // ComputeTableIndex
            thin[(size_t)idx].push_back(q_off);
        };
        for (const Range &loc : segs) {
            const int32_t from = loc.left, to = loc.right;
            if (word_size > to - from + 1) continue;
            const uint8_t *seq = Q + from;
            const uint8_t *target = seq + lut_width;
            int32_t offset;
            for (offset = from; offset <= to; offset++, seq++) {
                if (seq >= target) add_word(seq - lut_width, offset - lut_width);
                if (*seq & 0xfc) target = seq + lut_width + 1;
            }
            if (seq >= target) add_word(seq - lut_width, offset - lut_width);
        }
        int32_t longest = 0;
        for (const auto &ch : thin) longest = std::max(longest, (int32_t)ch.size());
        S->longest_chain = longest;
        // BlastSmallNaLookupTableNew fails when its 15-bit overflow array would not fit, and
        // LookupTableWrapInit then builds the standard table instead (core/lookup_wrap.c:126-135)
        bool standard = lut_type == 2;
        if (!standard) {
            int64_t need = 2;
            for (const auto &ch : thin) if (ch.size() > 1) need += (int64_t)ch.size() + 1;
            if (need >= 32768) standard = true;
        }
        if (standard) {
            // s_BlastNaLookupFinalize core/blast_nalookup.c:448-538: up to 3 hits in the cell, more in overflow
            b.lut_type = BN_LUT_NA;
            S->na_backbone.assign((size_t)(4 * b.hashsize), 0);
            for (int64_t i = 0; i < b.hashsize; i++) {
                const auto &ch = thin[(size_t)i];
                if (ch.empty()) continue;
                int32_t *cell = &S->na_backbone[(size_t)(4 * i)];
                cell[0] = (int32_t)ch.size();
                if (ch.size() <= 3) for (size_t j = 0; j < ch.size(); j++) cell[1 + j] = ch[j];
                else {
                    cell[1] = (int32_t)S->na_overflow.size();
                    S->na_overflow.insert(S->na_overflow.end(), ch.begin(), ch.end());
                }
            }
            if (S->na_overflow.empty()) S->na_overflow.push_back(0);
        } else {
            b.lut_type = BN_LUT_SMALL_NA;
            int64_t need = 2;
            for (const auto &ch : thin) {
                const int32_t n = (int32_t)ch.size();
                if (n > 1) need += n + 1;
            }
            S->backbone.assign((size_t)b.hashsize, (int16_t)-1);
            S->overflow.assign((size_t)need, (int16_t)0);
            int32_t cursor = 2;
            for (int64_t i = 0; i < b.hashsize; i++) {
                const auto &ch = thin[(size_t)i];
                if (ch.empty()) continue;
                if (ch.size() == 1) S->backbone[(size_t)i] = (int16_t)ch[0];
                else {
                    S->backbone[(size_t)i] = (int16_t)-cursor;
                    for (int32_t v : ch) S->overflow[(size_t)cursor++] = (int16_t)v;
                    S->overflow[(size_t)cursor++] = (int16_t)-1;
                }
            }
            b.overflow_len = cursor;
        }
    }
    if (!segs.empty() && word_size > lut_width && mask_at_hash) invert_locations(segs, concat_len, S->masked);

    // ---- assemble the batch -----------------------------------------------------------------------------
    b.query_start = S->query.data(); b.concat_len = concat_len;
    b.num_contexts = 2 * nq; b.contexts = S->ctx.data(); b.num_queries = nq;
    b.hashtable = S->hashtable.empty() ? nullptr : S->hashtable.data();
    b.next_pos = S->next_pos.empty() ? nullptr : S->next_pos.data();
    b.pv_array = S->pv.empty() ? nullptr : S->pv.data();
    b.backbone = S->backbone.empty() ? nullptr : S->backbone.data();
    b.overflow = S->overflow.empty() ? nullptr : S->overflow.data();
    b.na_backbone = S->na_backbone.empty() ? nullptr : S->na_backbone.data();
    b.na_overflow = S->na_overflow.empty() ? nullptr : S->na_overflow.data();
    b.na_overflow_len = (int64_t)S->na_overflow.size();
    b.masked_locations = S->masked.empty() ? nullptr : S->masked.data();
    b.n_masked_locations = (int32_t)(S->masked.size() / 2);
    for (const Range &r : segs) { S->segments.push_back(r.left); S->segments.push_back(r.right); }
    b.lookup_segments = S->segments.empty() ? nullptr : S->segments.data();
    b.n_lookup_segments = (int32_t)(S->segments.size() / 2);
    b.container_type = concat_len > 8000 ? BN_DIAG_HASH : BN_DIAG_ARRAY;   // kQueryLenForHashTable
    b.window_size = opt->window_size; b.scan_range = opt->scan_range;
    b.gap_algo = greedy ? BN_GAP_GREEDY : BN_GAP_DP;
    b.reward = reward; b.penalty = penalty; b.gap_open = gap_open; b.gap_extend = gap_extend;
    b.min_diag_separation = opt->min_diag_separation >= 0 ? opt->min_diag_separation : (mb ? 6 : 50);
    b.round_down = round_down ? 1 : 0;
    b.hsp_num_max = opt->hsp_num_max;
    b.percent_identity = opt->percent_identity; b.min_hit_length = opt->min_hit_length;
    b.hitlist_size = opt->hitlist_size ? opt->hitlist_size : 500;
    b.evalue_cutoff = evalue;
    b.low_score_perc = opt->low_score_perc >= 0 ? opt->low_score_perc : 0.15;
    *out = S;
    return BN_OK;
}

const BnQueryBatch *bn_setup_batch(const BnSetup *s) { return s ? &s->batch : nullptr; }
const double *bn_setup_kbp_std(const BnSetup *s) { return s ? s->kbp_std.data() : nullptr; }
const double *bn_setup_kbp_gap(const BnSetup *s) { return s ? s->kbp_gap.data() : nullptr; }
int32_t bn_setup_gap_x_dropoff_final(const BnSetup *s) { return s ? s->gap_x_dropoff_final : 0; }
int32_t bn_setup_longest_chain(const BnSetup *s) { return s ? s->longest_chain : 0; }
void bn_setup_free(BnSetup *s) { delete s; }

}  // extern "C"
