// pin.zig — scores the raw RGB8 pairs of make_pairs.py with fssimu2, exactly as oavif does at src/tq.zig:37.
// Build: see ../pin_fssimu2.md (zig 0.15.1).  SOURCE ONLY: there is no zig toolchain in the build image.
const std = @import("std");
const fssimu2 = @import("fssimu2");

const Case = struct { name: []const u8, w: u32, h: u32, kind: []const u8, seed: u64, strength: f64 };

pub fn main() !void {
    var gpa = std.heap.GeneralPurposeAllocator(.{}){};
    defer _ = gpa.deinit();
    const allocator = gpa.allocator();
    const args = try std.process.argsAlloc(allocator);
    defer std.process.argsFree(allocator, args);
    if (args.len < 2) return error.MissingIndexPath;
    const dir = std.fs.path.dirname(args[1]) orelse ".";
    const text = try std.fs.cwd().readFileAlloc(allocator, args[1], 1 << 20);
    defer allocator.free(text);
    const parsed = try std.json.parseFromSlice([]Case, allocator, text, .{});
    defer parsed.deinit();
    var out_buf: [4096]u8 = undefined;
    var stdout = std.fs.File.stdout().writer(&out_buf);
    const out = &stdout.interface;
    try out.writeAll("[\n");
    for (parsed.value, 0..) |c, i| {
        const n: usize = @as(usize, c.w) * c.h * 3;
        const src_path = try std.fmt.allocPrint(allocator, "{s}/{s}_src.rgb", .{ dir, c.name });
        defer allocator.free(src_path);
        const dst_path = try std.fmt.allocPrint(allocator, "{s}/{s}_dst.rgb", .{ dir, c.name });
        defer allocator.free(dst_path);
        const src = try std.fs.cwd().readFileAlloc(allocator, src_path, n);
        defer allocator.free(src);
        const dst = try std.fs.cwd().readFileAlloc(allocator, dst_path, n);
        defer allocator.free(dst);
        const score = try fssimu2.computeSsimu2(allocator, src, dst, c.w, c.h, 3, null); // tq.zig:37
        try out.print(" {{\"name\": \"{s}\", \"w\": {d}, \"h\": {d}, \"kind\": \"{s}\", \"seed\": {d}, \"strength\": {d}, \"score\": {d:.12}}}{s}\n", .{
            c.name, c.w, c.h, c.kind, c.seed, c.strength, score, if (i + 1 < parsed.value.len) "," else "",
        });
    }
    try out.writeAll("]\n");
    try out.flush();
}
