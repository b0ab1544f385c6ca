//! tq_batched.zig — batched (speculative) probing for oavif's target-quality search.  NEW FILE for src/.
//! Line-parallel with `findTargetQualityBatched` / `TQSearch` / `speculate` in oavif_b200/host/cpp/tq.hpp, which is
//! the compiled and tested form (tests/test_tq_policy.py, tests/test_host_search.py); this file is SOURCE ONLY
//! (no zig toolchain in the build image).  Zig 0.15.1.
//!
//! Contract: the search POLICY stays tq.zig's, byte for byte.  Speculative candidates — the q the policy wants now
//! and the qs it would want next under a few hypothetical outcomes — are encoded on host threads and scored in ONE
//! launch; then the sequential decision procedure is replayed using only scores of qs it asks for.  The chosen q,
//! the history and num_pass are exactly those of tq.findTargetQuality; probes it never asks for are `wasted` and
//! never influence a decision.
const std = @import("std");
const io = @import("io.zig");
const tq = @import("tq.zig");
const EncCtx = @import("main.zig").EncCtx;
const PassResult = tq.PassResult;

/// The per-pass decisions of tq.findTargetQuality (tq.zig:135-181) as a resumable state (tq.hpp: TQSearch).
const Search = struct {
    tgt: f64,
    tol: f64,
    max_pass: usize,
    history: std.ArrayList(PassResult),
    lo: u32 = 0,
    hi: u32 = 100,
    pass: usize = 0,
    done: bool = false,
    early: bool = false,

    fn next(s: *const Search, allocator: std.mem.Allocator) !?u32 {
        if (s.done or s.pass >= s.max_pass) return null;
        const q = if (s.pass == 0) tq.predictQFromScore(s.tgt) else try tq.interpolateQuantizer(allocator, s.lo, s.hi, s.history.items, s.tgt);
        for (s.history.items) |h| if (h.q == q) return null; // tq.zig:141-148
        return q;
    }

    fn record(s: *Search, allocator: std.mem.Allocator, q: u32, score: f64) !void {
        try s.history.append(allocator, .{ .q = q, .score = score });
        const abs_err = @abs(score - s.tgt);
        if (s.pass == 0) { // tq.zig:155-164
            const err_bound: u32 = @intFromFloat(@ceil(abs_err) * 4.0);
            if (score - s.tgt > 0) {
                s.hi = q;
                s.lo = if (q > err_bound) q - err_bound else 0;
            } else {
                s.lo = q;
                s.hi = @min(100, q + err_bound);
            }
        }
        if (abs_err < s.tol) { // tq.zig:167-168
            s.done = true;
            s.early = true;
            s.pass += 1;
            return;
        }
        if (s.pass > 0) { // tq.zig:171-176
            if (score > s.tgt) s.hi = q else s.lo = q;
        }
        if (s.lo >= s.hi -% 1) s.done = true; // tq.zig:179 (u32 arithmetic)
        s.pass += 1;
    }

    fn clone(s: *const Search, allocator: std.mem.Allocator) !Search {
        var c = s.*;
        c.history = try s.history.clone(allocator);
        return c;
    }
};

/// tq.hpp: speculate().  Hypothetical outcomes of the wanted probe, nearest misses first, alternating sign.
fn speculate(allocator: std.mem.Allocator, s: *const Search, width: usize, out: *std.ArrayList(u32)) !void {
    const want = (try s.next(allocator)) orelse return;
    try out.append(allocator, want);
    var i: usize = 0;
    while (i < 80 and out.items.len < width) : (i += 1) {
        const mag = s.tol + 0.25 + 0.5 * @as(f64, @floatFromInt(i / 2));
        const dlt = if (i & 1 == 1) -mag else mag;
        var h = try s.clone(allocator);
        defer h.history.deinit(allocator);
        try h.record(allocator, want, s.tgt + dlt);
        if (try h.next(allocator)) |nq| {
            if (std.mem.indexOfScalar(u32, out.items, nq) == null) try out.append(allocator, nq);
        }
    }
}

const Probe = struct { q: u32, score: f64 = 0, avif: std.ArrayListAligned(u8, null), planes: io.DecodedPlanes = undefined, staging: io.PlaneStaging = .{}, err: ?anyerror = null };

fn encodeDecodeOne(e: *const EncCtx, allocator: std.mem.Allocator, p: *Probe) void {
    var local = e.*; // encodeAvifToBuffer reads e.q / e.o / e.src only
    local.q = p.q;
    io.encodeAvifToBuffer(&local, allocator, &p.avif) catch |err| {
        p.err = err;
        return;
    };
    p.planes = io.decodeAvifPlanes(p.avif.items, &p.staging) catch |err| {
        p.err = err;
        return;
    };
}

pub fn findTargetQualityBatched(e: *EncCtx, allocator: std.mem.Allocator) !void {
    const o = &e.o;
    var s = Search{ .tgt = o.score_tgt, .tol = o.tolerance, .max_pass = o.max_pass, .history = try std.ArrayList(PassResult).initCapacity(allocator, 0) };
    defer s.history.deinit(allocator);
    var cache = try std.ArrayList(Probe).initCapacity(allocator, 0); // every probe made so far, asked-for or speculative
    defer {
        for (cache.items) |*p| {
            p.avif.deinit(allocator);
            p.staging.deinit();
        }
        cache.deinit(allocator);
    }

    while (try s.next(allocator)) |q| {
        var hit: ?usize = null;
        for (cache.items, 0..) |p, i| if (p.q == q) {
            hit = i;
        };
        if (hit == null) {
            var qs = try std.ArrayList(u32).initCapacity(allocator, 0);
            defer qs.deinit(allocator);
            try speculate(allocator, &s, o.batch, &qs);
            const first = cache.items.len;
            for (qs.items) |cq| {
                var known = false;
                for (cache.items) |p| if (p.q == cq) {
                    known = true;
                };
                if (!known) try cache.append(allocator, .{ .q = cq, .avif = try std.ArrayListAligned(u8, null).initCapacity(allocator, 0) });
            }
            const fresh = cache.items[first..];
            // encode + decode the fresh candidates concurrently (one host thread each), then ONE scoring launch
            var threads = try allocator.alloc(std.Thread, fresh.len);
            defer allocator.free(threads);
            for (fresh, 0..) |*p, i| threads[i] = try std.Thread.spawn(.{}, encodeDecodeOne, .{ e, allocator, p });
            for (threads) |t| t.join();
            for (fresh) |p| if (p.err) |err| return err;
            var planes = try allocator.alloc(io.DecodedPlanes, fresh.len);
            defer allocator.free(planes);
            const scores = try allocator.alloc(f64, fresh.len);
            defer allocator.free(scores);
            for (fresh, 0..) |p, i| planes[i] = p.planes;
            try e.scorer.?.scoreBatchYuv444(planes, scores);
            for (fresh, 0..) |*p, i| p.score = scores[i];
            for (cache.items, 0..) |p, i| if (p.q == q) {
                hit = i;
            };
        }
        const p = &cache.items[hit.?];
        // what computeScoreAtQuality does for a consumed pass (tq.zig:29-35): count it, keep its bytes as the cache
        e.t.num_pass += 1;
        if (e.buf.data) |*old| old.deinit(allocator);
        e.buf.data = try p.avif.clone(allocator);
        e.buf.q = q;
        e.buf.size = p.avif.items.len;
        e.q = q;
        e.t.score = p.score;
        try s.record(allocator, q, p.score);
    }
    if (s.early) return; // tq.zig:167-168: e.q / e.t.score stay those of the last pass

    // tq.zig:183-209, unchanged
    var best_q: ?u32 = null;
    var best_score: f64 = 0;
    var highest_q: u32 = 0;
    var highest_score: f64 = 0;
    for (s.history.items) |h| {
        if (h.score >= o.score_tgt and (best_q == null or h.q < best_q.?)) {
            best_q = h.q;
            best_score = h.score;
        }
        if (@max(h.score, 0) >= highest_score) {
            highest_score = h.score;
            highest_q = h.q;
        }
    }
    if (best_q) |q| {
        e.q = q;
        e.t.score = best_score;
        return;
    }
    e.q = highest_q;
    e.t.score = highest_score;
}
