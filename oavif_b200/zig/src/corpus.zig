//! corpus.zig — `oavif --corpus <images_dir> [--gpus G] [--workers-per-gpu W] out.csv`: scripts/measure.py in-process.
//! NEW FILE for src/.  Line-parallel with run_corpus / corpus_csv / corpus_summary in
//! oavif_b200/host/cpp/oavif_host.cpp (the compiled and tested form); SOURCE ONLY here.  Zig 0.15.1.
//!
//! What measure.py does per image — fork/exec oavif, wait, stat the output — becomes a call: G x W worker threads
//! (one scorer context each, bound to its GPU) pull image indices from ONE shared atomic counter; images are
//! independent, so there is no collective (no NCCL) and the only shared state is that counter.  The CSV has
//! measure.py's nine columns (measure.py:180-192) in image order, the summary its lines (measure.py:250-269).
const std = @import("std");
const io = @import("io.zig");
const tq = @import("tq.zig");
const a = @import("parse_args.zig");
const fssimu2 = @import("fssimu2");
const EncCtx = @import("main.zig").EncCtx;

const Row = struct {
    image: []const u8 = "",
    orig_bytes: u64 = 0,
    final_bytes: u64 = 0,
    ms: f64 = 0,
    passes: usize = 0,
    status: enum { ok, no_output, failed } = .failed,
    err: ?anyerror = null,
};

const Shared = struct {
    allocator: std.mem.Allocator,
    o: *const a.AvifEncOptions,
    dir: []const u8,
    files: []const []const u8,
    rows: []Row,
    next: std.atomic.Value(usize) = std.atomic.Value(usize).init(0),
};

fn worker(sh: *Shared, gpu: u8) void {
    var scorer: ?fssimu2.Scorer = null;
    defer if (scorer) |*s| s.deinit();
    var cap_w: u32 = 0;
    var cap_h: u32 = 0;
    var staging: io.PlaneStaging = .{};
    defer staging.deinit();
    while (true) {
        const i = sh.next.fetchAdd(1, .monotonic); // the shared work counter: whoever is free takes the next image
        if (i >= sh.files.len) return;
        const row = &sh.rows[i];
        row.image = sh.files[i];
        oneImage(sh, gpu, &scorer, &cap_w, &cap_h, &staging, row) catch |err| {
            row.status = .failed; // one bad image must not kill the sweep (measure.py:94-107)
            row.err = err;
        };
    }
}

fn oneImage(sh: *Shared, gpu: u8, scorer: *?fssimu2.Scorer, cap_w: *u32, cap_h: *u32, staging: *io.PlaneStaging, row: *Row) !void {
    const allocator = sh.allocator;
    const path = try std.fs.path.join(allocator, &.{ sh.dir, row.image });
    defer allocator.free(path);
    row.orig_bytes = (try std.fs.cwd().statFile(path)).size;
    var e: EncCtx = .{ .o = sh.o.* };
    e.src = try io.loadImage(allocator, path);
    defer e.src.deinit(allocator);
    e.rgb = if (e.src.channels == 3 and !e.src.hbd) e.src.data else try e.src.toRGB8(allocator);
    defer if (!(e.src.channels == 3 and !e.src.hbd)) allocator.free(e.rgb);
    e.w = @intCast(e.src.width);
    e.h = @intCast(e.src.height);
    if (scorer.* == null or e.w > cap_w.* or e.h > cap_h.*) { // one context per worker, regrown on demand
        if (scorer.*) |*s| s.deinit();
        cap_w.* = @max(cap_w.*, e.w);
        cap_h.* = @max(cap_h.*, e.h);
        scorer.* = try fssimu2.Scorer.init(gpu, cap_w.*, cap_h.*, sh.o.batch);
    }
    e.scorer = &scorer.*.?;
    e.staging = staging.*;
    defer staging.* = e.staging;
    var timer = try std.time.Timer.start(); // measure.py:61-64 times the whole oavif process
    try e.scorer.?.setSource(e.rgb, e.w, e.h);
    try tq.findTargetQuality(&e, allocator);
    defer e.buf.deinitCache(allocator);
    if (e.buf.q.? != e.q) { // main.zig:113: one more encode at the chosen q, not counted in num_pass
        var again = try std.ArrayListAligned(u8, null).initCapacity(allocator, 0);
        defer again.deinit(allocator);
        try io.encodeAvifToBuffer(&e, allocator, &again);
        e.buf.size = again.items.len;
    }
    row.ms = @as(f64, @floatFromInt(timer.read())) / 1e6;
    row.final_bytes = e.buf.size;
    row.passes = e.t.num_pass;
    row.status = if (e.buf.size > 0) .ok else .no_output;
}

fn humanBytes(buf: []u8, n: f64) []const u8 { // measure.py:31-38
    const units = [_][]const u8{ "B", "KiB", "MiB", "GiB", "TiB" };
    var size = n;
    for (units, 0..) |u, i| {
        if (size < 1024.0 or i == units.len - 1) return std.fmt.bufPrint(buf, "{d:.2} {s}", .{ size, u }) catch "";
        size /= 1024.0;
    }
    return "";
}

pub fn run(allocator: std.mem.Allocator, o: *const a.AvifEncOptions, dir_path: []const u8, csv_path: []const u8) !void {
    // measure.py:137-140: sorted .png / .jpg / .jpeg
    var names = try std.ArrayList([]const u8).initCapacity(allocator, 0);
    defer {
        for (names.items) |n| allocator.free(n);
        names.deinit(allocator);
    }
    var dir = try std.fs.cwd().openDir(dir_path, .{ .iterate = true });
    defer dir.close();
    var it = dir.iterate();
    while (try it.next()) |ent| {
        if (ent.kind != .file) continue;
        const ext = std.fs.path.extension(ent.name);
        if (std.ascii.eqlIgnoreCase(ext, ".png") or std.ascii.eqlIgnoreCase(ext, ".jpg") or std.ascii.eqlIgnoreCase(ext, ".jpeg"))
            try names.append(allocator, try allocator.dupe(u8, ent.name));
    }
    std.mem.sort([]const u8, names.items, {}, struct {
        fn lt(_: void, l: []const u8, r: []const u8) bool {
            return std.mem.lessThan(u8, l, r);
        }
    }.lt);
    if (names.items.len == 0) return error.NoImagesFound;

    const rows = try allocator.alloc(Row, names.items.len);
    defer allocator.free(rows);
    for (rows) |*r| r.* = .{};
    var sh = Shared{ .allocator = allocator, .o = o, .dir = dir_path, .files = names.items, .rows = rows };
    const cores = std.Thread.getCpuCount() catch 1;
    const wpg: usize = if (o.workers_per_gpu != 0) o.workers_per_gpu else @max(1, cores / o.gpus);
    const threads = try allocator.alloc(std.Thread, @as(usize, o.gpus) * wpg);
    defer allocator.free(threads);
    var wall = try std.time.Timer.start();
    for (threads, 0..) |*t, k| t.* = try std.Thread.spawn(.{}, worker, .{ &sh, @as(u8, @intCast(o.device + k / wpg)) });
    for (threads) |t| t.join();
    const wall_s = @as(f64, @floatFromInt(wall.read())) / 1e9;

    // CSV: measure.py:178-206
    const file = try std.fs.cwd().createFile(csv_path, .{});
    defer file.close();
    var wbuf: [8192]u8 = undefined;
    var fw = file.writer(&wbuf);
    const w = &fw.interface;
    try w.writeAll("Image,Original Bytes,Final Bytes,Savings Bytes,Savings %,Encoding Time (ms),Passes,Status,Error\r\n");
    var ok: usize = 0;
    var failed: usize = 0;
    var no_out: usize = 0;
    var orig_total: u64 = 0;
    var final_total: u64 = 0;
    var log_ratio: f64 = 0;
    var t_sum: f64 = 0;
    var p_sum: f64 = 0;
    var p_max: usize = 0;
    var p_min: usize = std.math.maxInt(usize);
    for (rows) |r| switch (r.status) {
        .ok => {
            const sav = if (r.orig_bytes > r.final_bytes) r.orig_bytes - r.final_bytes else 0;
            const pct = if (r.orig_bytes > 0) 100.0 * @as(f64, @floatFromInt(sav)) / @as(f64, @floatFromInt(r.orig_bytes)) else 0.0;
            try w.print("{s},{d},{d},{d},{d:.2},{d:.2},{d},ok,\r\n", .{ r.image, r.orig_bytes, r.final_bytes, sav, pct, r.ms, r.passes });
            ok += 1;
            orig_total += r.orig_bytes;
            final_total += r.final_bytes;
            if (r.orig_bytes > 0) log_ratio += @log(@as(f64, @floatFromInt(r.final_bytes)) / @as(f64, @floatFromInt(r.orig_bytes)));
            t_sum += r.ms;
            p_sum += @floatFromInt(r.passes);
            p_max = @max(p_max, r.passes);
            p_min = @min(p_min, r.passes);
        },
        .no_output => {
            try w.print("{s},{d},,,,{d:.2},{d},no-output,\r\n", .{ r.image, r.orig_bytes, r.ms, r.passes });
            no_out += 1;
        },
        .failed => {
            try w.print("{s},{d},,,,,,error,Error processing {s}: {s}\r\n", .{ r.image, r.orig_bytes, r.image, if (r.err) |e| @errorName(e) else "unknown" });
            failed += 1;
        },
    };
    try w.flush();

    // Summary: measure.py:250-269 (stddev / median lines are computed the same way in oavif_host.cpp: corpus_summary)
    var hb: [3][32]u8 = undefined;
    const print = std.debug.print;
    const savings = if (ok > 0 and orig_total > final_total) orig_total - final_total else 0;
    print("\nRun Summary\nImages: {d} ok, {d} no-output, {d} errors\n", .{ ok, no_out, failed });
    print("Total wall time: {d:.2} s\nThroughput: {d:.2} images/s\n", .{ wall_s, @as(f64, @floatFromInt(ok)) / wall_s });
    print("Input bytes throughput: {s}/s\n", .{humanBytes(&hb[0], @floor(@as(f64, @floatFromInt(orig_total)) / wall_s))});
    print("Output bytes throughput: {s}/s\n", .{humanBytes(&hb[1], @floor(@as(f64, @floatFromInt(final_total)) / wall_s))});
    print("\nCompression Totals\nOriginal total bytes: {d} ({s})\n", .{ orig_total, humanBytes(&hb[0], @floatFromInt(orig_total)) });
    print("Final total bytes:    {d} ({s})\n", .{ final_total, humanBytes(&hb[1], @floatFromInt(final_total)) });
    print("Savings (bytes):      {d} ({s})\n", .{ savings, humanBytes(&hb[2], @floatFromInt(savings)) });
    print("% saved (overall):    {d:.2}%\n", .{if (orig_total > 0) 100.0 * @as(f64, @floatFromInt(savings)) / @as(f64, @floatFromInt(orig_total)) else 0.0});
    if (ok > 0) {
        print("% saved (geometric mean across files): {d:.2}%\n", .{(1.0 - @exp(log_ratio / @as(f64, @floatFromInt(ok)))) * 100.0});
        print("\nTiming & Passes\nAverage encoding time: {d:.2} ms\n", .{t_sum / @as(f64, @floatFromInt(ok))});
        print("Average passes:        {d:.2} (max: {d}, min: {d})\n", .{ p_sum / @as(f64, @floatFromInt(ok)), p_max, p_min });
    }
    print("\nResults written to {s}\n", .{csv_path});
}
