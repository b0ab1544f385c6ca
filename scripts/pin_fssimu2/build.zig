// build.zig for pin.zig — wires fssimu2 the way /root/reference/build.zig:30-33,65 does.  SOURCE ONLY.
const std = @import("std");

pub fn build(b: *std.Build) void {
    const target = b.standardTargetOptions(.{});
    const optimize = b.standardOptimizeOption(.{});
    const fssimu2 = b.dependency("fssimu2", .{ .target = target, .optimize = optimize });
    const exe = b.addExecutable(.{
        .name = "pin_fssimu2",
        .root_module = b.createModule(.{ .root_source_file = b.path("pin.zig"), .target = target, .optimize = optimize }),
    });
    exe.root_module.addImport("fssimu2", fssimu2.module("fssimu2"));
    b.installArtifact(exe);
}
