//! Drop-in module named `fssimu2` for oavif: same entry point as the Zig package it replaces
//! (gianni-rosato/fssimu2 0.1.1, wired at build.zig:30-33,65 and called at src/tq.zig:37), backed by
//! the CUDA sm_100a library behind include/oavif_ssimu2.h.
//!
//! NOT COMPILED IN THIS REPO'S CI: the build container has no zig toolchain.  The file is kept
//! line-for-line parallel to oavif_b200/host/ssimu2.py (the ctypes mirror that the tests exercise),
//! so a reviewer can diff the two bindings.  Zig 0.15.1 syntax (build.zig.zon:5).
const std = @import("std");

pub const c = @cImport({
    @cInclude("oavif_ssimu2.h");
});

pub const Error = error{
    InvalidArgument, // OAVIF_SSIMU2_E_ARG
    CudaError, // OAVIF_SSIMU2_E_CUDA  (no device, launch failure: there is no CPU fallback)
    OutOfMemory, // OAVIF_SSIMU2_E_NOMEM
    InvalidState, // OAVIF_SSIMU2_E_STATE
    Unsupported, // OAVIF_SSIMU2_E_UNSUPPORTED
};

fn check(rc: c_int) Error!void {
    return switch (rc) {
        c.OAVIF_SSIMU2_OK => {},
        c.OAVIF_SSIMU2_E_ARG => Error.InvalidArgument,
        c.OAVIF_SSIMU2_E_NOMEM => Error.OutOfMemory,
        c.OAVIF_SSIMU2_E_STATE => Error.InvalidState,
        c.OAVIF_SSIMU2_E_UNSUPPORTED => Error.Unsupported,
        else => Error.CudaError,
    };
}

/// Same contract as the call at src/tq.zig:37:
///   `try fssimu2.computeSsimu2(allocator, e.rgb, decoded_rgb, e.w, e.h, 3, null)`
/// The allocator is unused (all device and pinned memory belongs to the library); the 7th argument is
/// always `null` in oavif and is ignored.
pub fn computeSsimu2(
    allocator: std.mem.Allocator,
    ref: []const u8,
    dist: []const u8,
    w: u32,
    h: u32,
    channels: u8,
    _: ?*anyopaque,
) Error!f64 {
    _ = allocator;
    const need: usize = @as(usize, w) * @as(usize, h) * @as(usize, channels);
    if (ref.len < need or dist.len < need) return Error.InvalidArgument;
    var score: f64 = 0;
    try check(c.oavif_ssimu2_compute_rgb8(ref.ptr, dist.ptr, w, h, channels, &score));
    return score;
}

/// Stateful form used by the patched search loop: the source pyramid is built once per image
/// (main.zig:86 `e.rgb`), every pass uploads only the candidate (tq.zig:150).
pub const Scorer = struct {
    ctx: *c.oavif_ssimu2_ctx,
    max_batch: u32,

    pub fn init(device: c_int, max_w: u32, max_h: u32, max_batch: u32) Error!Scorer {
        var ctx: ?*c.oavif_ssimu2_ctx = null;
        try check(c.oavif_ssimu2_ctx_create(device, max_w, max_h, max_batch, &ctx));
        return .{ .ctx = ctx.?, .max_batch = max_batch };
    }

    pub fn deinit(self: *Scorer) void {
        c.oavif_ssimu2_ctx_destroy(self.ctx);
        self.* = undefined;
    }

    pub fn setSource(self: *Scorer, rgb: []const u8, w: u32, h: u32) Error!void {
        try check(c.oavif_ssimu2_set_source_rgb8(self.ctx, rgb.ptr, w, h, @as(usize, w) * 3));
    }

    pub fn scoreRgb8(self: *Scorer, dist: []const u8, w: u32) Error!f64 {
        var score: f64 = 0;
        try check(c.oavif_ssimu2_score_rgb8(self.ctx, dist.ptr, @as(usize, w) * 3, &score));
        return score;
    }

    /// Decoded planes straight from `decoder.*.image` (io.zig:463): no avifImageYUVToRGB, no repack.
    /// `rgba_path` = the decoded image carries an alpha plane (io.zig:473).
    pub fn scoreYuv444(
        self: *Scorer,
        planes: [3][*]const u8,
        row_bytes: [3]u32,
        depth: u32,
        matrix_coefficients: u16,
        rgba_path: bool,
    ) Error!f64 {
        var score: f64 = 0;
        try check(c.oavif_ssimu2_score_yuv444(self.ctx, planes[0], planes[1], planes[2], row_bytes[0], row_bytes[1], row_bytes[2], @intCast(depth), @intCast(matrix_coefficients), @intFromBool(rgba_path), &score));
        return score;
    }

    /// Image.toRGB8 on the device (io.zig:57-133): the loaders' native layouts, 1..4 channels, 8- or 16-bit.
    pub fn setSourcePixels(self: *Scorer, pixels: []const u8, w: u32, h: u32, channels: u8, hbd: bool) Error!void {
        const bps: usize = if (hbd) 2 else 1;
        try check(c.oavif_ssimu2_set_source_pixels(self.ctx, pixels.ptr, w, h, @as(usize, w) * channels * bps, channels, if (hbd) 16 else 8));
    }

    /// Readings of the published algorithm this library carries side by side until fssimu2 0.1.1 vectors settle them
    /// (include/oavif_ssimu2.h: OAVIF_SSIMU2_OPT_*; scripts/pin_fssimu2/which_variant.py names the matching one).
    pub const Option = enum(c_int) { blur = 1, weights = 2, tile_path = 3, source_rows = 4, transfer = 5, vertical_order = 6 };

    pub fn setOption(self: *Scorer, option: Option, value: c_int) Error!void {
        try check(c.oavif_ssimu2_set_option(self.ctx, @intFromEnum(option), value));
    }

    /// The sample array encodeAvifToBuffer rebuilds from the source in every pass (io.zig:566-609: 8 -> 10 bit
    /// (v*1023+127)/255, 16 -> 10 bit v >> 6, 16 -> 8 bit v >> 8), made ONCE from the pixels setSourcePixels staged.
    /// `out` holds w*h*channels samples: u16 for depth 10, u8 for depth 8 (pass it as bytes).
    pub fn sourceSamples(self: *Scorer, out_depth: u8, out: []u8) Error!void {
        try check(c.oavif_ssimu2_source_samples(self.ctx, out_depth, out.ptr, out.len));
    }

    /// Batched probing (tq_batched.zig): n decoded candidates (io.DecodedPlanes layout: same depth, strides and
    /// matrix for all of them) against the cached source in one pass over the device.
    pub fn scoreBatchYuv444(self: *Scorer, cands: anytype, scores: []f64) Error!void {
        const n = cands.len;
        if (n == 0 or n > self.max_batch or n > 16 or scores.len < n) return Error.InvalidArgument;
        var y: [16]?*const anyopaque = undefined;
        var u: [16]?*const anyopaque = undefined;
        var v: [16]?*const anyopaque = undefined;
        for (cands, 0..) |p, i| {
            y[i] = p.planes[0];
            u[i] = p.planes[1];
            v[i] = p.planes[2];
        }
        const p0 = cands[0];
        try check(c.oavif_ssimu2_score_batch_yuv444(self.ctx, @intCast(n), &y, &u, &v, p0.row_bytes[0], p0.row_bytes[1], p0.row_bytes[2], @intCast(p0.depth), @intCast(p0.matrix_coefficients), @intFromBool(p0.has_alpha), scores.ptr));
    }

    /// Pipelined form: submit returns at once (the upload runs on the context's copy stream under the previous
    /// submission's kernels), wait retires the oldest submission.  Up to two in flight; planes must stay valid
    /// until the matching wait.
    pub fn submitYuv444(self: *Scorer, planes: [3][*]const u8, row_bytes: [3]u32, depth: u32, matrix_coefficients: u16, rgba_path: bool) Error!void {
        const y: [1]?*const anyopaque = .{planes[0]};
        const u: [1]?*const anyopaque = .{planes[1]};
        const v: [1]?*const anyopaque = .{planes[2]};
        try check(c.oavif_ssimu2_submit_yuv444(self.ctx, 1, &y, &u, &v, row_bytes[0], row_bytes[1], row_bytes[2], @intCast(depth), @intCast(matrix_coefficients), @intFromBool(rgba_path)));
    }

    pub fn wait(self: *Scorer) Error!f64 {
        var score: f64 = 0;
        try check(c.oavif_ssimu2_wait(self.ctx, &score));
        return score;
    }

    /// Batched probing (new in tq.zig): n candidate decodes of one image in one pass over the device.
    pub fn scoreBatchRgb8(self: *Scorer, dists: []const [*]const u8, w: u32, scores: []f64) Error!void {
        if (dists.len == 0 or dists.len > self.max_batch or scores.len < dists.len) return Error.InvalidArgument;
        try check(c.oavif_ssimu2_score_batch_rgb8(self.ctx, @intCast(dists.len), @ptrCast(dists.ptr), @as(usize, w) * 3, scores.ptr));
    }

    pub fn lastError(self: *const Scorer) [*:0]const u8 {
        return c.oavif_ssimu2_last_error(self.ctx);
    }
};

/// Pinned staging for io.zig's decode buffers (cudaHostAlloc): H2D becomes a DMA.
pub fn pinnedAlloc(bytes: usize) Error![]u8 {
    const p = c.oavif_ssimu2_pinned_alloc(bytes) orelse return Error.OutOfMemory;
    return @as([*]u8, @ptrCast(p))[0..bytes];
}

pub fn pinnedFree(buf: []u8) void {
    c.oavif_ssimu2_pinned_free(buf.ptr);
}
