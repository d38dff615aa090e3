// video_interface.hpp -- VI register decode and scan-out sequencing.
//
// Role mirrors parallel-rdp/video_interface.{hpp,cpp}: VI registers -> per-scanline sampling parameters, then
// the scan-out stages (extract -> AA/dither-filter fetch -> divot -> scale/gamma/field blend) on the GPU.
#pragma once
#include "rdp_b200_backend.h"
#include "rdp_types.h"
#include <stdint.h>
#include <vector>

namespace RDP
{
enum class VIRegister
{
	Control = 0,
	Origin,
	Width,
	Intr,
	VCurrentLine,
	Timing,
	VSync,
	HSync,
	Leap,
	HStart,
	VStart,
	VBurst,
	XScale,
	YScale,
	Count
};

enum VIControlFlagBits
{
	VI_CONTROL_TYPE_BLANK_BIT = 0 << 0,
	VI_CONTROL_TYPE_RESERVED_BIT = 1 << 0,
	VI_CONTROL_TYPE_RGBA5551_BIT = 2 << 0,
	VI_CONTROL_TYPE_RGBA8888_BIT = 3 << 0,
	VI_CONTROL_TYPE_MASK = 3 << 0,
	VI_CONTROL_GAMMA_DITHER_ENABLE_BIT = 1 << 2,
	VI_CONTROL_GAMMA_ENABLE_BIT = 1 << 3,
	VI_CONTROL_DIVOT_ENABLE_BIT = 1 << 4,
	VI_CONTROL_SERRATE_BIT = 1 << 6,
	VI_CONTROL_AA_MODE_RESAMP_EXTRA_ALWAYS_BIT = 0 << 8,
	VI_CONTROL_AA_MODE_RESAMP_EXTRA_BIT = 1 << 8,
	VI_CONTROL_AA_MODE_RESAMP_ONLY_BIT = 2 << 8,
	VI_CONTROL_AA_MODE_RESAMP_REPLICATE_BIT = 3 << 8,
	VI_CONTROL_AA_MODE_MASK = 3 << 8,
	VI_CONTROL_DITHER_FILTER_ENABLE_BIT = 1 << 16,
	VI_CONTROL_META_AA_BIT = 1 << 17,
	VI_CONTROL_META_SCALE_BIT = 1 << 18
};

static inline uint32_t make_vi_start_register(uint32_t start_value, uint32_t end_value)
{
	return ((start_value & 0x3ff) << 16) | (end_value & 0x3ff);
}

static inline uint32_t make_vi_scale_register(uint32_t scale_factor, uint32_t bias)
{
	return ((bias & 0xfff) << 16) | (scale_factor & 0xfff);
}

constexpr uint32_t VI_V_SYNC_NTSC = 525;
constexpr uint32_t VI_V_SYNC_PAL = 625;
constexpr uint32_t VI_H_OFFSET_NTSC = 108;
constexpr uint32_t VI_H_OFFSET_PAL = 128;
constexpr uint32_t VI_V_OFFSET_NTSC = 34;
constexpr uint32_t VI_V_OFFSET_PAL = 44;
constexpr uint32_t VI_V_RES_NTSC = 480;
constexpr uint32_t VI_V_RES_PAL = 576;
constexpr int VI_SCANOUT_WIDTH = 640;
constexpr uint32_t VI_V_RES_MAX = VI_V_RES_PAL;
constexpr uint32_t VI_V_END_PAL = (VI_V_OFFSET_PAL + VI_V_RES_PAL) | 1;
constexpr uint32_t VI_V_END_NTSC = (VI_V_OFFSET_NTSC + VI_V_RES_NTSC) | 1;
constexpr uint32_t VI_V_END_MAX = VI_V_END_PAL;
constexpr uint32_t VI_MAX_OUTPUT_SCANLINES = VI_V_RES_MAX >> 1;

struct ScanoutOptions
{
	unsigned crop_overscan_pixels = 0;
	struct CropRect
	{
		unsigned left = 0, right = 0, top = 0, bottom = 0;
		bool enable = false;
	} crop_rect;
	unsigned downscale_steps = 0;
	bool persist_frame_on_invalid_input = false;
	// Previous frame persists for one extra frame (needed for weave interlacing; matches the CPU oracle).
	bool blend_previous_frame = false;
	// Bob de-interlacing when true, weave when false.
	bool upscale_deinterlacing = true;
	struct
	{
		bool aa = true;
		bool scale = true;
		bool serrate = true;
		bool dither_filter = true;
		bool divot_filter = true;
		bool gamma_dither = true;
	} vi;
};

// Replaces Vulkan::ImageHandle: RGBA8 pixels in device memory, valid until the next scanout().
struct ScanoutImage
{
	const void *device_rgba8 = nullptr;
	unsigned width = 0;
	unsigned height = 0;
	explicit operator bool() const { return device_rgba8 != nullptr; }
};

// Replaces {Vulkan::BufferHandle, Vulkan::Fence}.
struct VIScanoutBuffer
{
	std::vector<uint32_t> pixels;
	unsigned width = 0;
	unsigned height = 0;
};

class VideoInterface
{
public:
	enum PerScanlineRegisterBits
	{
		PER_SCANLINE_HSTART_BIT = 1 << 0,
		PER_SCANLINE_XSCALE_BIT = 1 << 1
	};
	using PerScanlineRegisterFlags = uint32_t;

	// Everything the emulator writes: the register file is snapshotted per scan-out so a scan-out can execute
	// later (on the command thread) while the emulator already programs the next frame's registers.
	struct PerScanlineRegisterState
	{
		uint32_t latched_state;
		uint32_t line_state[VI_V_END_MAX];
	};
	struct RegisterFile
	{
		uint32_t vi_registers[unsigned(VIRegister::Count)] = {};
		struct
		{
			PerScanlineRegisterState h_start;
			PerScanlineRegisterState x_scale;
			PerScanlineRegisterFlags flags = 0;
			unsigned line = 0;
			bool ended = false;
		} per_line_state;
	};

	void init(rdpb_device *device, size_t rdram_size);
	// Internal upscaling: scan out of the upscaled domain at (1 << log2) times the resolution (ref video_interface.cpp:1215-1317).
	void set_upscaling_log2(uint32_t log2) { scale_log2 = log2; }
	void set_vi_register(VIRegister reg, uint32_t value);
	// Copies the register file for one scan-out and resets the per-scanline latches (as scanout() always did).
	// The registers as capture_registers() would hand them out, without consuming the per-scanline state.
	void peek_registers(RegisterFile &out) const { out = regfile; }
	// RDRAM byte range a scan-out with these registers can read; false when nothing would be scanned out.
	bool scanout_memory_range(const RegisterFile &file, uint32_t &offset, uint32_t &length) const;
	void capture_registers(RegisterFile &out);
	ScanoutImage scanout(const ScanoutOptions &options, const RegisterFile &file);
	ScanoutImage scanout(const ScanoutOptions &options);

	void begin_vi_register_per_scanline(PerScanlineRegisterFlags flags);
	void set_vi_register_for_scanline(PerScanlineRegisterBits reg, uint32_t value);
	void latch_vi_register_for_scanline(unsigned vi_line);
	void end_vi_register_per_scanline();

private:
	rdpb_device *device = nullptr;
	size_t rdram_size = 0;
	uint32_t scale_log2 = 0;
	RegisterFile regfile;

	uint32_t frame_count = 0;
	uint32_t last_valid_frame_count = 0;
	bool previous_frame_blank = false;
	bool have_prev_scanout = false;
	unsigned prev_width = 0, prev_height = 0;               // full-size image the next field merges with
	unsigned prev_result_width = 0, prev_result_height = 0; // what scanout() returned (after downscale steps)

	rdpb::ViRegs decode_vi_registers(const RegisterFile &file, rdpb::ViLine *lines) const;
	static void vram_range(const rdpb::ViRegs &regs, bool divot, int64_t &offset, int64_t &length);
	void clear_per_scanline_state();
};
}
