"""The C++ restatement of oavif's search policy (oavif_b200/host/cpp/tq.hpp) against an independent
Python transcription of /root/reference/src/tq.zig made for this test, on known answers, randomised
score curves, and the batched mode's "same decisions as sequential" guarantee.  CPU only."""
import math
import random

import numpy as np
import pytest

from oavif_b200.host import harness as H


# ---- transcription of tq.zig (line numbers in comments) ---------------------------------------------
def z_round(x):  # Zig @round: half away from zero
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def z_predict(tgt):  # tq.zig:40-43
    return int(min(100.0, z_round(6.83 * math.exp(0.0282 * tgt))))


def z_linear(s, q, t):  # tq.zig:45-51
    if len(s) < 2 or s[1] == s[0]:
        return None
    return q[0] + (q[1] - q[0]) * ((t - s[0]) / (s[1] - s[0]))


def z_quadratic(s, q, t):  # tq.zig:53-71
    if len(s) < 3:
        return None
    x0, x1, x2 = s[0], s[1], s[2]
    y0, y1, y2 = q[0], q[1], q[2]
    den = (x0 - x1) * (x0 - x2) * (x1 - x2)
    if abs(den) < 0.001:
        return None
    a = (x2 * (y1 - y0) + x1 * (y0 - y2) + x0 * (y2 - y1)) / den
    b = (x2 * x2 * (y0 - y1) + x1 * x1 * (y2 - y0) + x0 * x0 * (y1 - y2)) / den
    c = (x1 * x2 * (x1 - x2) * y0 + x2 * x0 * (x2 - x0) * y1 + x0 * x1 * (x0 - x1) * y2) / den
    return a * t * t + b * t + c


def z_interp(lo, hi, hist, t):  # tq.zig:73-122
    bs = (lo + hi) // 2
    if not hist:
        return bs
    srt = sorted(hist, key=lambda p: p[1])
    s = [p[1] for p in srt]
    q = [float(p[0]) for p in srt]
    clampq = lambda r: int(min(100.0, max(0.0, z_round(r))))
    if len(hist) == 1:
        pred = bs
    elif len(hist) == 2:
        r = z_linear(s, q, t)
        pred = clampq(r) if r is not None else bs
    else:
        r = z_quadratic(s, q, t)
        if r is not None:
            pred = clampq(r)
        else:
            r = z_linear(s, q, t)
            pred = clampq(r) if r is not None else bs
    return min(max(pred, lo), hi)


def z_search(score_of_q, tgt=80.0, tol=2.0, max_pass=6):  # tq.zig:124-210
    hist, lo, hi = [], 0, 100
    for p in range(max_pass):
        q = z_predict(tgt) if p == 0 else z_interp(lo, hi, hist, tgt)
        if any(h[0] == q for h in hist):
            break
        sc = score_of_q(q)
        hist.append((q, sc))
        err = abs(sc - tgt)
        if p == 0:
            eb = int(math.ceil(err) * 4.0)
            if sc - tgt > 0:
                hi, lo = q, (q - eb if q > eb else 0)
            else:
                lo, hi = q, min(100, q + eb)
        if err < tol:
            return q, sc, hist, True
        if p > 0:
            if sc > tgt:
                hi = q
            else:
                lo = q
        if lo >= hi - 1:
            break
    best = None
    hq, hs = 0, 0.0
    for q, sc in hist:
        if sc >= tgt and (best is None or q < best[0]):
            best = (q, sc)
        if max(sc, 0.0) >= hs:
            hs, hq = sc, q
    return (best[0], best[1], hist, False) if best else (hq, hs, hist, False)


# ---- known answers --------------------------------------------------------------------------------------
def test_predict_q_table():
    # SURVEY.md Appendix B, probed from the formula at tq.zig:40-43
    table = {30: 16, 40: 21, 50: 28, 60: 37, 70: 49, 75: 57, 80: 65, 85: 75, 90: 86, 95: 100, 100: 100}
    for t, q in table.items():
        assert H.predict_q(float(t)) == q == z_predict(float(t))


def test_interpolation_cases():
    assert H.interpolate_q(10, 90, [], 80.0) == 50                                  # no history: bisect
    assert H.interpolate_q(10, 91, [(65, 70.0)], 80.0) == 50                        # one point: bisect (floor)
    assert H.interpolate_q(0, 100, [(65, 70.0), (85, 90.0)], 80.0) == 75            # inverse linear
    assert H.interpolate_q(0, 100, [(85, 90.0), (65, 70.0)], 80.0) == 75            # order independent (sorted by score)
    assert H.interpolate_q(0, 100, [(65, 70.0), (85, 70.0)], 80.0) == 50            # equal scores -> bisect
    assert H.interpolate_q(70, 72, [(65, 70.0), (85, 90.0)], 80.0) == 72            # clamp to bounds
    assert H.interpolate_q(0, 100, [(40, 60.0), (60, 75.0), (80, 85.0)], 80.0) == z_interp(0, 100, [(40, 60.0), (60, 75.0), (80, 85.0)], 80.0)
    # quadratic uses the three LOWEST scores (Appendix D #4), not the three nearest the target
    h4 = [(30, 50.0), (50, 68.0), (70, 79.0), (90, 93.0)]
    assert H.interpolate_q(0, 100, h4, 80.0) == z_interp(0, 100, h4, 80.0)
    assert H.interpolate_q(0, 100, h4, 80.0) == z_interp(0, 100, h4[:3], 80.0)
    # nearly collinear scores: denominator < 0.001 -> falls back to linear on the two lowest
    h3 = [(50, 70.0), (51, 70.01), (52, 70.02)]
    assert H.interpolate_q(0, 100, h3, 80.0) == z_interp(0, 100, h3, 80.0)


def test_search_traces_on_reference_shaped_curves():
    curves = {
        "linear": lambda q: 40 + 0.5 * q,
        "saturating": lambda q: 100 - 60 * math.exp(-q / 35.0),
        "steep": lambda q: 20 + 0.9 * q,
        "flat": lambda q: 79.0 + 0.001 * q,
        "never": lambda q: 10 + 0.3 * q,
        "always": lambda q: 95.0 + 0.01 * q,
        "negative": lambda q: -50 + 1.0 * q,
    }
    for name, f in curves.items():
        for tgt, tol, mp in ((80.0, 2.0, 6), (90.0, 1.0, 12), (60.0, 5.0, 3), (80.0, 1.0, 1)):
            want = z_search(f, tgt, tol, mp)
            got = H.tq_search(f, tgt, tol, mp)
            assert (got.q, got.num_pass, bool(got.early_exit)) == (want[0], len(want[2]), want[3]), (name, tgt)
            assert got.score == want[1]
            assert got.history() == want[2]


def test_early_exit_accepts_below_target_within_tolerance():
    # Appendix D #5: first probe lands 1.5 below the target with tolerance 2 -> returned as is
    r = H.tq_search(lambda q: 78.5, 80.0, 2.0, 6)
    assert (r.q, r.score, r.num_pass, r.early_exit) == (65, 78.5, 1, 1)


def test_no_pass_meets_target_picks_highest_score_with_tie_to_later_entry():
    f = {65: 50.0}.get
    seq = []

    def score(q):
        seq.append(q)
        return 50.0 if len(seq) != 2 else 50.0

    r = H.tq_search(score, 80.0, 2.0, 6)
    want = z_search(lambda q: 50.0, 80.0, 2.0, 6)
    assert (r.q, r.score, r.num_pass) == (want[0], want[1], len(want[2]))


@pytest.mark.parametrize("seed", range(40))
def test_randomised_curves_sequential_and_batched(seed):
    rng = random.Random(seed)
    a, b, c = rng.uniform(-20, 70), rng.uniform(0.1, 1.2), rng.uniform(0.0, 0.004)
    noise = {q: rng.gauss(0, rng.choice([0.0, 0.3, 2.0])) for q in range(101)}
    f = lambda q: a + b * q - c * q * q + noise[q]
    tgt = rng.choice([50.0, 70.0, 80.0, 90.0, 95.0])
    tol = rng.choice([1.0, 2.0, 4.0])
    mp = rng.choice([1, 2, 4, 6, 12])
    want = z_search(f, tgt, tol, mp)
    got = H.tq_search(f, tgt, tol, mp)
    assert (got.q, got.score, got.history(), bool(got.early_exit)) == (want[0], want[1], want[2], want[3])
    for width in (2, 4, 8):
        calls = []

        def fb(q):
            calls.append(q)
            return f(q)

        bat = H.tq_search_batched(fb, width, tgt, tol, mp)
        # same chosen q, same score, same history and pass accounting as the sequential procedure
        assert (bat.q, bat.score, bat.history(), bat.num_pass) == (got.q, got.score, got.history(), got.num_pass)
        assert bat.probes == len(calls) and len(set(calls)) == len(calls)          # no q probed twice
        assert bat.wasted == bat.probes - bat.num_pass
        assert bat.device_passes <= bat.num_pass                                   # never more round trips


# ---- margin report (tq.hpp decisionMargins) -----------------------------------------------------------------------
def test_decision_margins_bracket_every_flip():
    """Moving a pass's score by less than its margin never changes the search; moving it by just more does."""
    rng = np.random.default_rng(5)
    checked = 0
    for trial in range(30):
        a, b = rng.uniform(0.15, 0.5), rng.uniform(40, 75)
        curve = lambda q: min(99.0, b + a * q + 0.002 * q * q)                     # monotone score(q)
        tgt, tol = float(rng.choice([70, 80, 85, 90])), float(rng.choice([1.0, 2.0]))
        r = H.tq_search(curve, tgt, tol, 6)
        hist = r.history()
        margins = H.tq_margins(hist, tgt, tol, 6)
        assert len(margins) == len(hist)

        def replay(i, delta):   # what the search asks for after pass i when that score moves, and its final pick
            moved = [(q, s + (delta if j == i else 0.0)) for j, (q, s) in enumerate(hist)]
            nxt = H.tq_search(lambda q: dict(moved)[q] if q in dict(moved) else curve(q), tgt, tol, 6)
            return [q for q, _ in nxt.history()][: i + 2], nxt.early_exit if len(nxt.history()) == i + 1 else None

        for i, (up, dn) in enumerate(margins):
            assert up > 0 and dn > 0
            for sign, m in ((1, up), (-1, dn)):
                if m >= 4.0:
                    continue
                base = replay(i, 0.0)
                assert replay(i, sign * m * 0.98) [0][: i + 1] == base[0][: i + 1]
                inside, outside = replay(i, sign * m * 0.98), replay(i, sign * m * 1.02)
                if inside == base and outside != base:
                    checked += 1
    assert checked > 20     # most finite margins are flips of the NEXT quantizer / the stop, which this replay sees


def test_decision_margins_known_cases():
    # one pass inside the tolerance: the exit flips when |score - tgt| reaches the tolerance
    (up, dn), = H.tq_margins([(65, 80.5)], 80.0, 2.0, 6)
    assert abs(up - 1.5) < 1e-6          # 80.5 + 1.5 = 82.0 -> no longer < tolerance
    assert abs(dn - 2.5) < 1e-6          # 80.5 - 2.5 = 78.0
    # first pass outside the tolerance: the sign and ceil(|err|) * 4 decide the bracket (tq.zig:155-164)
    (up, dn), = H.tq_margins([(65, 84.3)], 80.0, 2.0, 1)
    assert up >= 4.0 and abs(dn - 2.3) < 1e-6    # max_pass 1: nothing follows, only the tolerance exit can change
