// tq.hpp — executable restatement of oavif's target-quality search policy
// (/root/reference/src/tq.zig) for the C++ host harness.  Policy only: how a score is obtained is the
// caller's business (`Probe`: encode at q, decode, score — tq.zig:21-38).  Every decision is written
// so that the f64 operation order equals the Zig source (tq.zig:40-122, 124-210), because
// interpolateQuantizer rounds a continuous function of past scores and "same quantizer" depends on it.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <optional>
#include <vector>

namespace oavif_host {

struct PassResult {  // tq.zig:16-19
    uint32_t q;
    double score;
};

struct TQOptions {  // the fields of AvifEncOptions the search reads (parse_args.zig:49-63)
    double score_tgt = 80.0;
    double tolerance = 2.0;
    uint32_t max_pass = 6;
};

struct TQResult {
    uint32_t q = 0;          // e.q on return
    double score = 0.0;      // e.t.score on return
    uint32_t num_pass = 0;   // e.t.num_pass
    std::vector<PassResult> history;
    bool early_exit = false; // returned through the tolerance test (tq.zig:167-168)
};

// tq.zig:40-43
inline uint32_t predictQFromScore(double tgt)
{
    const double q = 6.83 * std::exp(0.0282 * tgt);
    return (uint32_t)std::min(100.0, std::round(q));
}

// tq.zig:45-51
inline std::optional<double> linearInterpolate(const std::vector<double> &scores, const std::vector<double> &qualities,
                                               double target)
{
    if (scores.size() < 2) return std::nullopt;
    if (scores[1] == scores[0]) return std::nullopt;
    const double t = (target - scores[0]) / (scores[1] - scores[0]);
    return qualities[0] + (qualities[1] - qualities[0]) * t;
}

// tq.zig:53-71
inline std::optional<double> quadraticInterpolate(const std::vector<double> &scores,
                                                  const std::vector<double> &qualities, double target)
{
    if (scores.size() < 3) return std::nullopt;
    const double x0 = scores[0], x1 = scores[1], x2 = scores[2];
    const double y0 = qualities[0], y1 = qualities[1], y2 = qualities[2];
    const double denom = (x0 - x1) * (x0 - x2) * (x1 - x2);
    if (std::fabs(denom) < 0.001) return std::nullopt;
    const double coeff_a = (x2 * (y1 - y0) + x1 * (y0 - y2) + x0 * (y2 - y1)) / denom;
    const double coeff_b = (x2 * x2 * (y0 - y1) + x1 * x1 * (y2 - y0) + x0 * x0 * (y1 - y2)) / denom;
    const double coeff_c = (x1 * x2 * (x1 - x2) * y0 + x2 * x0 * (x2 - x0) * y1 + x0 * x1 * (x0 - x1) * y2) / denom;
    return coeff_a * target * target + coeff_b * target + coeff_c;
}

inline uint32_t roundClampQ(double r)
{  // @intFromFloat(std.math.clamp(@round(r), 0, 100)) — tq.zig:109,114,116
    return (uint32_t)std::min(100.0, std::max(0.0, std::round(r)));
}

// tq.zig:73-122
inline uint32_t interpolateQuantizer(uint32_t lo_bound, uint32_t hi_bound, const std::vector<PassResult> &history,
                                     double target)
{
    const uint32_t binary_search = (lo_bound + hi_bound) / 2;
    if (history.empty()) return binary_search;
    std::vector<PassResult> sorted(history);
    // std.mem.sort is an insertion-friendly block sort, stable; stable_sort keeps equal scores in order
    std::stable_sort(sorted.begin(), sorted.end(),
                     [](const PassResult &l, const PassResult &r) { return l.score < r.score; });
    std::vector<double> scores, qualities;
    for (const auto &p : sorted) {
        scores.push_back(p.score);
        qualities.push_back((double)p.q);
    }
    uint32_t pred;
    switch (history.size()) {
    case 1: pred = binary_search; break;
    case 2:
        if (auto r = linearInterpolate(scores, qualities, target)) pred = roundClampQ(*r);
        else pred = binary_search;
        break;
    default:
        if (auto r = quadraticInterpolate(scores, qualities, target)) pred = roundClampQ(*r);
        else if (auto lr = linearInterpolate(scores, qualities, target)) pred = roundClampQ(*lr);
        else pred = binary_search;
        break;
    }
    return std::min(std::max(pred, lo_bound), hi_bound);  // std.math.clamp(pred, lo, hi)
}

// The per-pass decision procedure of findTargetQuality (tq.zig:124-181) as a resumable state
// machine, so that the sequential loop and the batched (speculative) loop share one copy of it.
class TQSearch {
  public:
    explicit TQSearch(const TQOptions &o) : o_(o) {}

    // q the next pass wants, or nullopt when the loop is over (max_pass reached, range collapsed,
    // tolerance hit, or the wanted q was already probed — tq.zig:141-148).
    std::optional<uint32_t> next() const
    {
        if (done_ || pass_ >= o_.max_pass) return std::nullopt;
        const uint32_t q = pass_ == 0 ? predictQFromScore(o_.score_tgt)
                                      : interpolateQuantizer(lo_, hi_, history_, o_.score_tgt);
        for (const auto &h : history_)
            if (h.q == q) return std::nullopt;
        return q;
    }

    // Feed the score of the pass that probed `q` (tq.zig:150-180).
    void record(uint32_t q, double score)
    {
        history_.push_back({q, score});
        last_q_ = q;
        last_score_ = score;
        const double abs_err = std::fabs(score - o_.score_tgt);
        if (pass_ == 0) {
            const uint32_t err_bound = (uint32_t)(std::ceil(abs_err) * 4.0);
            if (score - o_.score_tgt > 0) {
                hi_ = q;
                lo_ = q > err_bound ? q - err_bound : 0;
            } else {
                lo_ = q;
                hi_ = std::min<uint32_t>(100, q + err_bound);
            }
        }
        if (abs_err < o_.tolerance) {
            done_ = true;
            early_exit_ = true;
            ++pass_;
            return;
        }
        if (pass_ > 0) {
            if (score > o_.score_tgt) hi_ = q;
            else lo_ = q;
        }
        if (lo_ >= hi_ - 1) done_ = true;  // u32 arithmetic as in the reference (tq.zig:179)
        ++pass_;
    }

    // tq.zig:167-168 (early return keeps the current q) and 183-209 (post-loop choice).
    TQResult finish() const
    {
        TQResult r;
        r.history = history_;
        r.num_pass = (uint32_t)history_.size();
        r.early_exit = early_exit_;
        if (early_exit_) {
            r.q = last_q_;
            r.score = last_score_;
            return r;
        }
        std::optional<uint32_t> best_q;
        double best_score = 0, highest_score = 0;
        uint32_t highest_q = 0;
        for (const auto &h : history_) {
            if (h.score >= o_.score_tgt && (!best_q || h.q < *best_q)) {
                best_q = h.q;
                best_score = h.score;
            }
            if (std::max(h.score, 0.0) >= highest_score) {
                highest_score = h.score;
                highest_q = h.q;
            }
        }
        if (best_q) {
            r.q = *best_q;
            r.score = best_score;
        } else {
            r.q = highest_q;
            r.score = highest_score;
        }
        return r;
    }

    uint32_t lo() const { return lo_; }
    uint32_t hi() const { return hi_; }
    bool early_exit() const { return early_exit_; }
    const std::vector<PassResult> &history() const { return history_; }

  private:
    TQOptions o_;
    std::vector<PassResult> history_;
    uint32_t lo_ = 0, hi_ = 100, pass_ = 0;
    uint32_t last_q_ = 0;
    double last_score_ = 0.0;
    bool done_ = false, early_exit_ = false;
};

// tq.zig:124-210, sequential: one probe per pass.
inline TQResult findTargetQuality(const TQOptions &o, const std::function<double(uint32_t)> &probe)
{
    TQSearch s(o);
    while (auto q = s.next()) s.record(*q, probe(*q));
    return s.finish();
}

// ---- margin report ------------------------------------------------------------------------------------------
// How far each score of a finished search was from changing what the search did next.  Every decision of
// tq.zig:135-209 that reads a score is covered at once by replaying the procedure with that one score moved:
//   * the pass's own outcome — sign of score - target (tq.zig:157, 172), ceil(|err|) * 4 (tq.zig:156), the
//     tolerance exit (tq.zig:167), the collapse test (tq.zig:179) and, through the next predicted quantizer,
//     every @round / clamp of interpolateQuantizer (tq.zig:109-121) and the already-probed break (tq.zig:141-148);
//   * the post-loop choice `score >= target` / highest score (tq.zig:189-196).
// flip_up / flip_down are the smallest increase / decrease of the pass's score (searched up to `limit`, resolved
// by bisection to 1e-9) after which the search would have asked for a different next quantizer, stopped or
// continued differently, or returned a different final q; `limit` itself means "no flip within the limit".
// A scorer that agrees with the reference's to better than min(flip_up, flip_down) of every pass provably
// reproduces the whole search: same probes, same q, same bytes.
struct DecisionMargin {
    uint32_t pass, q;
    double score, flip_up, flip_down;
};

inline std::vector<DecisionMargin> decisionMargins(const TQOptions &o, const std::vector<PassResult> &history,
                                                   double limit = 4.0)
{
    struct Key {
        bool has_next, early;
        uint32_t next_q, final_q;
        bool operator==(const Key &k) const
        {
            return has_next == k.has_next && early == k.early && next_q == k.next_q && final_q == k.final_q;
        }
    };
    // outcome of the search when pass i's score is moved by delta and nothing else changes
    auto outcome = [&](size_t i, double delta) {
        TQSearch s(o);
        for (size_t j = 0; j <= i; ++j) s.record(history[j].q, history[j].score + (j == i ? delta : 0.0));
        const auto nq = s.next();
        Key k{nq.has_value(), s.early_exit(), nq ? *nq : 0u, 0u};
        // the final choice with the REAL later passes appended (it only reads scores): stays comparable as long
        // as the next quantizer is unchanged, and a changed one is a flip anyway
        TQSearch f(o);
        for (size_t j = 0; j < history.size(); ++j) {
            f.record(history[j].q, history[j].score + (j == i ? delta : 0.0));
            if (f.early_exit()) break;
        }
        k.final_q = f.finish().q;
        return k;
    };
    std::vector<DecisionMargin> out;
    for (size_t i = 0; i < history.size(); ++i) {
        const Key ref = outcome(i, 0.0);
        DecisionMargin m{(uint32_t)i, history[i].q, history[i].score, limit, limit};
        for (int sign = -1; sign <= 1; sign += 2) {
            double lo = 0.0, hi = -1.0;
            for (double d = 1e-7; d <= limit * 1.0000001; d *= 1.25) {   // the outcome is piecewise constant, not monotone
                if (!(outcome(i, sign * d) == ref)) {
                    hi = d;
                    break;
                }
                lo = d;
            }
            if (hi < 0.0) continue;
            for (int it = 0; it < 60 && hi - lo > 1e-9; ++it) {
                const double mid = 0.5 * (lo + hi);
                if (outcome(i, sign * mid) == ref) lo = mid;
                else hi = mid;
            }
            (sign > 0 ? m.flip_up : m.flip_down) = hi;
        }
        out.push_back(m);
    }
    return out;
}

inline double minMargin(const std::vector<DecisionMargin> &ms, double limit = 4.0)
{
    double m = limit;
    for (const auto &d : ms) m = std::min(m, std::min(d.flip_up, d.flip_down));
    return m;
}

// Batched mode (new, additive).  `probe_batch` scores several quantizers in one device pass.  The
// candidates are speculative: the q the policy wants now plus the qs it would want next under a few
// hypothetical outcomes.  Decisions are then REPLAYED through the unmodified sequential procedure
// using only scores of qs that procedure asks for, so the chosen q, the history and num_pass are
// exactly those of the sequential run; speculative probes that the policy never asks for are counted
// separately (`wasted`) and never influence a decision.
struct BatchedStats {
    uint32_t device_passes = 0, probes = 0, wasted = 0;
};

// score(q) is monotone and, over the range a search visits, close to linear with a slope the history reveals
// (two probes) or that is typical of libaom at speed 9 (one probe: ~0.45 points per q step).  `expected_error`
// is where that prior puts (score - target) of the probe the policy wants next.
inline double expectedError(const TQSearch &s, const TQOptions &o, uint32_t want)
{
    const auto &h = s.history();
    if (h.empty()) return 0.0;   // pass 0 probes predictQFromScore(target): the prior says "on target"
    double slope = 0.45;
    if (h.size() >= 2) {
        const PassResult &a = h[h.size() - 2], &b = h.back();
        if (a.q != b.q) slope = std::min(2.0, std::max(0.05, (b.score - a.score) / ((double)b.q - (double)a.q)));
    }
    return h.back().score + slope * ((double)want - (double)h.back().q) - o.score_tgt;
}

// The q the policy wants now, then the qs it would want NEXT under hypothetical outcomes of that probe, most
// plausible first: outcomes are ranked by their distance from the monotone prior's expectation, so a narrow batch
// spends its extra encodes where the search is most likely to go.
inline std::vector<uint32_t> speculate(const TQSearch &s, const TQOptions &o, uint32_t width)
{
    std::vector<uint32_t> qs;
    auto want = s.next();
    if (!want) return qs;
    qs.push_back(*want);
    if (width <= 1) return qs;
    // Hypothetical errors: on the first pass the next q only depends on ceil(|err|) and the sign (tq.zig:155-164),
    // so half-unit steps hit every bucket; on later passes they sample the interpolation densely enough.  An
    // outcome inside the tolerance ends the search and needs no candidate.
    std::vector<double> deltas;
    for (int i = 0; i < 80; ++i) {
        const double mag = o.tolerance + 0.25 + 0.5 * (i / 2);
        deltas.push_back((i & 1) ? -mag : mag);
    }
    const double centre = expectedError(s, o, *want);
    std::stable_sort(deltas.begin(), deltas.end(),
                     [&](double a, double b) { return std::fabs(a - centre) < std::fabs(b - centre); });
    for (double dlt : deltas) {
        if (qs.size() >= width) break;
        TQSearch h = s;
        h.record(*want, o.score_tgt + dlt);
        if (auto nq = h.next())
            if (std::find(qs.begin(), qs.end(), *nq) == qs.end()) qs.push_back(*nq);
    }
    return qs;
}

inline TQResult findTargetQualityBatched(const TQOptions &o, uint32_t width,
                                         const std::function<std::vector<double>(const std::vector<uint32_t> &)> &probe_batch,
                                         BatchedStats *stats = nullptr)
{
    TQSearch s(o);
    std::vector<PassResult> cache;  // every score obtained so far, asked-for or speculative
    BatchedStats st;
    while (auto q = s.next()) {
        auto hit = std::find_if(cache.begin(), cache.end(), [&](const PassResult &p) { return p.q == *q; });
        if (hit == cache.end()) {
            std::vector<uint32_t> qs = speculate(s, o, std::max<uint32_t>(1, width));
            qs.erase(std::remove_if(qs.begin(), qs.end(), [&](uint32_t c) {
                         return std::any_of(cache.begin(), cache.end(), [&](const PassResult &p) { return p.q == c; });
                     }), qs.end());
            const std::vector<double> sc = probe_batch(qs);
            ++st.device_passes;
            st.probes += (uint32_t)qs.size();
            for (size_t i = 0; i < qs.size(); ++i) cache.push_back({qs[i], sc[i]});
            hit = std::find_if(cache.begin(), cache.end(), [&](const PassResult &p) { return p.q == *q; });
        }
        s.record(*q, hit->score);
    }
    TQResult r = s.finish();
    st.wasted = st.probes - r.num_pass;
    if (stats) *stats = st;
    return r;
}

}  // namespace oavif_host
