// Touzet & Varré p-value -> score threshold, host side of the library.  Replaces pvalue2score and its helpers
//   min_score_range / round_pwm / best_score / worst_score / create_Q / find_largest_alpha     inference/_h2_Touzet.jl:1-187
// which the reference runs in Julia per effective PWM segment (<= 15 columns) before every filtered scan.  Float64 throughout and
// the same order of additions as the reference's Dict-of-sorted-keys walk (keys ascending, equal keys accumulated score-major,
// base-minor), so results are bit-identical to the Python mirror (inference.pvalue2score) that the tests compare against.
// No device work: the DP has a few thousand states; the point of having it here is that a Julia / C host gets the thresholds from
// the same library call sequence as the scans, and that 900 PWMs cost milliseconds instead of seconds of interpreter time.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <limits>
#include <utility>

extern "C" int32_t mb200_pvalue2score(mb200_ctx* ctx, const double* pwm, int32_t m, double pval, double eps, const double* bg,
                                      double* score, int32_t* found) {
    // ctx may be NULL (pure host arithmetic: no device, no error text)
    if (!pwm || !bg || !score || !found || m < 1) MB_FAIL(ctx, MB200_E_INVALID, "pvalue2score: bad arguments");
    if (!(pval >= 0.0 && pval <= 1.0)) MB_FAIL(ctx, MB200_E_INVALID, "pvalue must be in [0,1]");
    // columns by decreasing (max - min), stable (min_score_range); entries rounded down to the granularity (round_pwm)
    std::vector<int> order(m);
    std::vector<double> delta(m);
    for (int j = 0; j < m; ++j) {
        double mx = pwm[j], mn = pwm[j];
        for (int a = 1; a < 4; ++a) { mx = std::max(mx, pwm[a * m + j]); mn = std::min(mn, pwm[a * m + j]); }
        delta[j] = mx - mn; order[j] = j;
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return -delta[x] < -delta[y]; });
    std::vector<double> pe((size_t)4 * m), colmax(m), colmin(m);
    for (int j = 0; j < m; ++j) {
        for (int a = 0; a < 4; ++a) pe[(size_t)a * m + j] = std::floor(pwm[a * m + order[j]] / eps) * eps;
        colmax[j] = colmin[j] = pe[j];
        for (int a = 1; a < 4; ++a) { colmax[j] = std::max(colmax[j], pe[(size_t)a * m + j]); colmin[j] = std::min(colmin[j], pe[(size_t)a * m + j]); }
    }
    auto suffix = [&](const std::vector<double>& v, int i) { double s = 0.0; for (int j = i; j < m; ++j) s += v[j]; return s; };
    const double alpha = suffix(colmin, 0);
    const double inf = std::numeric_limits<double>::infinity();
    std::vector<double> keys(1, 0.0), vals(1, 1.0);
    std::vector<std::pair<double, double>> tw;
    for (int i = 0; i < m; ++i) {
        const double bs = i + 1 < m ? suffix(colmax, i + 1) : 0.0, ws = i + 1 < m ? suffix(colmin, i + 1) : 0.0;
        tw.clear();
        for (size_t k = 0; k < keys.size(); ++k)
            for (int a = 0; a < 4; ++a) {
                const double t = keys[k] + pe[(size_t)a * m + i];
                if (alpha - bs <= t && t <= inf - ws) tw.emplace_back(t, vals[k] * bg[a]);
            }
        std::stable_sort(tw.begin(), tw.end(), [](const std::pair<double, double>& x, const std::pair<double, double>& y) { return x.first < y.first; });
        keys.clear(); vals.clear();
        for (size_t j = 0; j < tw.size(); ++j) {
            if (keys.empty() || tw[j].first != keys.back()) { keys.push_back(tw[j].first); vals.push_back(0.0); }
            vals.back() += tw[j].second;                           // Q[i][t] += ..., in encounter order
        }
    }
    double q_sum = 0.0;
    for (double v : vals) q_sum += v;
    *found = 0; *score = 0.0;
    for (size_t k = 0; k < keys.size(); ++k) {                     // find_largest_alpha
        if (q_sum >= pval) { *found = 1; *score = keys[k]; }
        else { *found = 1; *score = keys[k]; return MB200_OK; }
        q_sum -= vals[k];
    }
    return MB200_OK;
}
