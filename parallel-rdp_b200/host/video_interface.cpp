// video_interface.cpp -- VI register decode and scan-out sequencing; see video_interface.hpp.
//
// decode_vi_registers() and the first part of scanout() follow the reference's VideoInterface::decode_vi_registers /
// scanout (parallel-rdp/video_interface.cpp:177-330, Copyright (c) 2020 Themaister, MIT licence -- notice in
// video_interface.hpp) closely, statement for statement in places: the VI's register arithmetic (start / end clamps,
// PAL / NTSC geometry, the guard band, per-line scaling) is a list of hardware facts with one right answer, and the
// host side of this project stays C++ in the reference's shape by design.  Everything downstream of the decoded
// per-line parameters (kernels, staging, streams) is this project's own.
//
// Reference behaviour re-implemented here: parallel-rdp/video_interface.cpp
//   register decode   :177-367       per-scanline registers :1126-1213
//   scan-out control  :1215-1426     scale pass geometry    :719-999
#include "video_interface.hpp"
#include "rdp_vi.cuh"
#include <algorithm>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

using namespace rdpb;

namespace RDP
{
void VideoInterface::init(rdpb_device *device_, size_t rdram_size_)
{
	device = device_;
	rdram_size = rdram_size_;
}

void VideoInterface::set_vi_register(VIRegister reg, uint32_t value)
{
	regfile.vi_registers[unsigned(reg)] = value;
}

// ---------------------------------------------------------------------------------------------
// Per-scanline HStart / XScale (raster effects)
// ---------------------------------------------------------------------------------------------
void VideoInterface::begin_vi_register_per_scanline(PerScanlineRegisterFlags flags)
{
	regfile.per_line_state.flags = flags;
	regfile.per_line_state.h_start.latched_state = regfile.vi_registers[unsigned(VIRegister::HStart)];
	regfile.per_line_state.x_scale.latched_state = regfile.vi_registers[unsigned(VIRegister::XScale)];
	regfile.per_line_state.h_start.line_state[0] = regfile.per_line_state.h_start.latched_state;
	regfile.per_line_state.x_scale.line_state[0] = regfile.per_line_state.x_scale.latched_state;
	regfile.per_line_state.line = 0;
	regfile.per_line_state.ended = false;
}

void VideoInterface::set_vi_register_for_scanline(PerScanlineRegisterBits reg, uint32_t value)
{
	if ((regfile.per_line_state.flags & reg) == 0)
	{
		fprintf(stderr, "[rdp-b200] VI register %u was not flagged in begin_vi_register_per_scanline, ignoring.\n", unsigned(reg));
		return;
	}
	if (reg == PER_SCANLINE_HSTART_BIT)
		regfile.per_line_state.h_start.latched_state = value;
	else if (reg == PER_SCANLINE_XSCALE_BIT)
		regfile.per_line_state.x_scale.latched_state = value;
}

void VideoInterface::latch_vi_register_for_scanline(unsigned vi_line)
{
	vi_line = std::min(vi_line, VI_V_END_MAX - 1);
	if (vi_line <= regfile.per_line_state.line)
	{
		fprintf(stderr, "[rdp-b200] vi_line %u is not monotonically increasing (current %u), ignoring.\n", vi_line, regfile.per_line_state.line);
		return;
	}
	// Lines up to the latch point keep the previous values.
	const unsigned from = regfile.per_line_state.line;
	for (unsigned l = from + 1; l < vi_line; l++)
	{
		regfile.per_line_state.h_start.line_state[l] = regfile.per_line_state.h_start.line_state[from];
		regfile.per_line_state.x_scale.line_state[l] = regfile.per_line_state.x_scale.line_state[from];
	}
	regfile.per_line_state.h_start.line_state[vi_line] = regfile.per_line_state.h_start.latched_state;
	regfile.per_line_state.x_scale.line_state[vi_line] = regfile.per_line_state.x_scale.latched_state;
	regfile.per_line_state.line = vi_line;
}

void VideoInterface::end_vi_register_per_scanline()
{
	if (regfile.per_line_state.flags == 0 || regfile.per_line_state.ended)
		return;
	const unsigned from = regfile.per_line_state.line;
	for (unsigned l = from + 1; l < VI_V_END_MAX; l++)
	{
		regfile.per_line_state.h_start.line_state[l] = regfile.per_line_state.h_start.line_state[from];
		regfile.per_line_state.x_scale.line_state[l] = regfile.per_line_state.x_scale.line_state[from];
	}
	regfile.per_line_state.ended = true;
}

void VideoInterface::clear_per_scanline_state()
{
	regfile.per_line_state.flags = 0;
	regfile.per_line_state.ended = false;
}

// ---------------------------------------------------------------------------------------------
// Register decode: VI registers -> bounding boxes + per-output-line sampling parameters
// ---------------------------------------------------------------------------------------------
ViRegs VideoInterface::decode_vi_registers(const RegisterFile &file, ViLine *lines) const
{
	ViRegs reg = {};
	const auto R = [&](VIRegister r) { return file.vi_registers[unsigned(r)]; };

	reg.status = R(VIRegister::Control);
	reg.vi_width = int(R(VIRegister::Width) & 0xfff);
	reg.vi_offset = int(R(VIRegister::Origin) & 0xffffff);
	reg.v_current_line = int(R(VIRegister::VCurrentLine) & 1);
	reg.v_start = int((R(VIRegister::VStart) >> 16) & 0x3ff);

	int y_start = int((R(VIRegister::YScale) >> 16) & 0xfff);
	const int y_add = int(R(VIRegister::YScale) & 0xfff);
	reg.init_y_add = y_add;

	int v_end = int(R(VIRegister::VStart) & 0x3ff);
	const int v_sync = int(R(VIRegister::VSync) & 0x3ff);
	reg.is_pal = unsigned(v_sync) > (VI_V_SYNC_NTSC + 25) ? 1 : 0;

	const int v_end_max = int(reg.is_pal ? VI_V_END_PAL : VI_V_END_NTSC);
	v_end = std::min(v_end, v_end_max);
	reg.v_start = std::min(reg.v_start, v_end_max);

	const int v_offset = int(reg.is_pal ? VI_V_OFFSET_PAL : VI_V_OFFSET_NTSC);
	const int h_offset = int(reg.is_pal ? VI_H_OFFSET_PAL : VI_H_OFFSET_NTSC);
	reg.v_res = (v_end - reg.v_start) >> 1;
	reg.v_start = (reg.v_start - v_offset) / 2;
	if (reg.v_start < 0)
	{
		y_start -= y_add * reg.v_start;
		reg.v_start = 0;
	}
	reg.v_res = std::min(reg.v_res, int(VI_MAX_OUTPUT_SCANLINES) - reg.v_start);

	int h_start_clamp_lo = INT_MAX, h_end_clamp_hi = 0, h_start_lo = INT_MAX, h_end_hi = 0;
	reg.max_x = 0;
	const bool degenerate_y = reg.v_res <= 0;

	if (lines)
	{
		const ViLine zero = {};
		for (int l = 0; l < int(VI_MAX_OUTPUT_SCANLINES); l++)
			if (degenerate_y || l < reg.v_start || l >= reg.v_start + reg.v_res)
				lines[l] = zero;
	}

	const auto analyze = [&](int x_start, int x_add, int h_start, int h_end, ViLine *out) {
		bool left_clamp = false, right_clamp = false;
		h_start -= h_offset;
		h_end -= h_offset;
		if (h_start < 0)
		{
			x_start -= x_add * h_start;
			h_start = 0;
			left_clamp = true;
		}
		if (h_end > VI_SCANOUT_WIDTH)
		{
			h_end = VI_SCANOUT_WIDTH;
			right_clamp = true;
		}
		// 8 / 7 pixel guard band unless the active area was clamped to the screen edge.
		const int h_start_clamp = h_start + (left_clamp ? 0 : 8);
		const int h_end_clamp = h_end - (right_clamp ? 0 : 7);
		const int h_res = h_end - h_start;
		reg.max_x = std::max(reg.max_x, (x_start + h_res * x_add) >> 10);
		h_start_lo = std::min(h_start_lo, h_start);
		h_end_hi = std::max(h_end_hi, h_end);
		h_start_clamp_lo = std::min(h_start_clamp_lo, h_start_clamp);
		h_end_clamp_hi = std::max(h_end_clamp_hi, h_end_clamp);
		if (out)
		{
			out->h_start = h_start;
			out->h_start_clamp = h_start_clamp;
			out->h_end_clamp = h_end_clamp;
			out->x_start = x_start;
			out->x_add = x_add;
			out->y_start = y_start;
			out->y_add = y_add;
			out->y_base = 0;
		}
	};

	if (degenerate_y || !file.per_line_state.ended)
	{
		ViLine info = {};
		analyze(int((R(VIRegister::XScale) >> 16) & 0xfff), int(R(VIRegister::XScale) & 0xfff),
		        int((R(VIRegister::HStart) >> 16) & 0x3ff), int(R(VIRegister::HStart) & 0x3ff), &info);
		if (lines)
			for (int l = reg.v_start; l < reg.v_start + reg.v_res; l++)
				lines[l] = info;
	}
	else
	{
		for (int l = reg.v_start; l < reg.v_start + reg.v_res; l++)
		{
			const int effective = 2 * l + v_offset + int(reg.v_current_line == 0);
			const uint32_t xs = file.per_line_state.x_scale.line_state[effective];
			const uint32_t hs = file.per_line_state.h_start.line_state[effective];
			analyze(int((xs >> 16) & 0xfff), int(xs & 0xfff), int((hs >> 16) & 0x3ff), int(hs & 0x3ff), lines ? &lines[l] : nullptr);
		}
	}

	reg.max_y = (y_start + reg.v_res * y_add) >> 10;
	reg.h_start = h_start_lo;
	reg.h_res = h_end_hi - h_start_lo;
	reg.h_start_clamp = h_start_clamp_lo;
	reg.h_res_clamp = h_end_clamp_hi - h_start_clamp_lo;
	return reg;
}

// ---------------------------------------------------------------------------------------------
// Scan-out
// ---------------------------------------------------------------------------------------------
static bool shifted_rect(int32_t *dst, int x, int y, int w, int h, int dx, int dy)
{
	// Rectangle shifted by (-crop_left, -crop_top) and clipped to the positive quadrant.
	int x0 = x + dx, y0 = y + dy, x1 = x + dx + w, y1 = y + dy + h;
	x0 = std::max(x0, 0);
	y0 = std::max(y0, 0);
	if (x1 <= x0 || y1 <= y0)
		return false;
	dst[0] = x0;
	dst[1] = y0;
	dst[2] = x1;
	dst[3] = y1;
	return true;
}

void VideoInterface::capture_registers(RegisterFile &out)
{
	if (regfile.per_line_state.flags != 0)
		out = regfile;
	else
	{
		// No raster effects programmed: the 14 registers are the whole story.
		memcpy(out.vi_registers, regfile.vi_registers, sizeof(out.vi_registers));
		out.per_line_state.flags = 0;
		out.per_line_state.ended = false;
	}
	clear_per_scanline_state();
}

// RDRAM bytes a scan-out with these registers can read (ref scanout_memory_range, video_interface.cpp:369-396): the
// displayed rows with a 2-pixel border for the AA / dither filters plus one more column for the divot.
void VideoInterface::vram_range(const ViRegs &regs, bool divot, int64_t &offset, int64_t &length)
{
	const int pixel_size_log2 = (regs.status & VI_CONTROL_TYPE_MASK) == VI_CONTROL_TYPE_RGBA8888_BIT ? 2 : 1;
	const int aa_width = regs.max_x + 2 + 4 + int(divot) * 2;
	const int aa_height = regs.max_y + 1 + 4;
	offset = regs.vi_offset & ~((1 << pixel_size_log2) - 1);
	offset += int64_t((divot ? -3 : -2) + -2 * regs.vi_width) << pixel_size_log2;
	length = (int64_t(aa_height) * regs.vi_width + aa_width) << pixel_size_log2;
}

bool VideoInterface::scanout_memory_range(const RegisterFile &file, uint32_t &offset, uint32_t &length) const
{
	ViLine lines[VI_MAX_OUTPUT_SCANLINES];
	const ViRegs regs = decode_vi_registers(file, lines);
	if (regs.vi_offset == 0 || (regs.status & VI_CONTROL_TYPE_RGBA5551_BIT) == 0)
		return false;
	int64_t o = 0, l = 0;
	vram_range(regs, true, o, l);
	if (l <= 0)
		return false;
	// Clamp to RDRAM (a range that runs off either end is read modulo the RDRAM size by the kernels: be generous).
	if (o < 0 || o + l > int64_t(rdram_size))
	{
		offset = 0;
		length = uint32_t(rdram_size);
		return true;
	}
	offset = uint32_t(o);
	length = uint32_t(l);
	return true;
}

ScanoutImage VideoInterface::scanout(const ScanoutOptions &options)
{
	RegisterFile file;
	capture_registers(file);
	return scanout(options, file);
}

ScanoutImage VideoInterface::scanout(const ScanoutOptions &options, const RegisterFile &file)
{
	ScanoutImage result;
	ViLine lines[VI_MAX_OUTPUT_SCANLINES];
	ViRegs regs = decode_vi_registers(file, lines);

	if (regs.vi_offset == 0)
	{
		have_prev_scanout = false;
		return result;
	}

	if (!options.vi.serrate)
		regs.status &= ~uint32_t(VI_CONTROL_SERRATE_BIT);

	bool is_aa = (regs.status & VI_CONTROL_AA_MODE_MASK) < VI_CONTROL_AA_MODE_RESAMP_ONLY_BIT;
	bool is_bilinear = (regs.status & VI_CONTROL_AA_MODE_MASK) < VI_CONTROL_AA_MODE_RESAMP_REPLICATE_BIT;
	is_aa = is_aa && options.vi.aa;
	is_bilinear = is_bilinear && options.vi.scale;
	regs.status &= ~uint32_t(VI_CONTROL_AA_MODE_MASK | VI_CONTROL_META_AA_BIT | VI_CONTROL_META_SCALE_BIT);
	if (is_aa)
		regs.status |= VI_CONTROL_META_AA_BIT;
	if (is_bilinear)
		regs.status |= VI_CONTROL_META_SCALE_BIT;
	if (!options.vi.gamma_dither)
		regs.status &= ~uint32_t(VI_CONTROL_GAMMA_DITHER_ENABLE_BIT);
	if (!options.vi.divot_filter)
		regs.status &= ~uint32_t(VI_CONTROL_DIVOT_ENABLE_BIT);
	if (!options.vi.dither_filter)
		regs.status &= ~uint32_t(VI_CONTROL_DITHER_FILTER_ENABLE_BIT);

	const bool is_blank = (regs.status & VI_CONTROL_TYPE_RGBA5551_BIT) == 0;
	if (is_blank && previous_frame_blank)
	{
		frame_count++;
		have_prev_scanout = false;
		return result;
	}
	if (is_blank)
		have_prev_scanout = false;
	regs.status |= VI_CONTROL_TYPE_RGBA5551_BIT;
	previous_frame_blank = is_blank;

	if (regs.h_res <= 0 || regs.h_start >= VI_SCANOUT_WIDTH)
	{
		frame_count++;
		if (options.persist_frame_on_invalid_input && (frame_count - last_valid_frame_count < 4) && have_prev_scanout)
		{
			result.device_rgba8 = rdpb_vi_output_device_ptr(device);
			result.width = prev_result_width;
			result.height = prev_result_height;
		}
		else
			have_prev_scanout = false;
		return result;
	}
	last_valid_frame_count = frame_count;

	const bool degenerate = regs.h_res <= 0 || regs.v_res <= 0;
	const bool divot = (regs.status & VI_CONTROL_DIVOT_ENABLE_BIT) != 0;
	const bool serrate = (regs.status & VI_CONTROL_SERRATE_BIT) != 0 && !options.upscale_deinterlacing;

	// Internal upscaling: everything measured in framebuffer or output pixels scales with the image
	// (ref video_interface.cpp:1299-1317); the per-pixel steps (x_add, y_add) do not.
	const int scale = 1 << scale_log2;
	// RDRAM range this scan-out can read (ref scanout_memory_range, video_interface.cpp:369-396): a 2-pixel border for
	// the AA / dither filters plus one more column for the divot.
	uint32_t vram_lo = 0, vram_hi = 0;
	{
		const int pixel_size_log2 = (regs.status & VI_CONTROL_TYPE_MASK) == VI_CONTROL_TYPE_RGBA8888_BIT ? 2 : 1;
		int64_t offset = 0, length = 0;
		vram_range(regs, divot, offset, length);
		if (offset >= 0 && length > 0 && offset + length <= int64_t(rdram_size))
		{
			vram_lo = uint32_t(offset);
			vram_hi = uint32_t(offset + length);
		}
		if (scale > 1)
		{
			// The framebuffer may hold pixels the RDP never rendered (CPU-drawn screens): bring the upscaled domain up
			// to date over that range (ref vram_fetch_stage, video_interface.cpp:412-420).
			FbInfo range = {};
			range.color_addr = range.depth_addr = uint32_t(offset) & uint32_t(rdram_size - 1);
			range.width = uint32_t(length >> pixel_size_log2);
			range.height = 1;
			range.fmt = pixel_size_log2 == 2 ? uint32_t(FB_RGBA8888) : uint32_t(FB_RGBA5551);
			range.rdram_size = uint32_t(rdram_size);
			rdpb_launch_upscale_sync(device, &range, RDPB_UPSCALE_SYNC_PRE);
		}
	}
	if (scale > 1)
	{
		regs.h_start *= scale;
		regs.h_start_clamp *= scale;
		regs.v_start *= scale;
		regs.h_res *= scale;
		regs.h_res_clamp *= scale;
		regs.v_res *= scale;
		regs.max_x = regs.max_x * scale + (scale - 1);
		regs.max_y = regs.max_y * scale + (scale - 1);
		for (auto &line : lines)
		{
			line.h_start *= scale;
			line.h_start_clamp *= scale;
			line.h_end_clamp *= scale;
			line.x_start *= scale;
			line.y_start *= scale;
			line.y_base *= scale;
		}
	}

	ViFrame f = {};
	// NB: blank (0) / reserved (1) types were promoted to 5551 / 8888 above; the first blank frame after a
	// valid one is still scanned out with that promoted type, exactly like the oracle does.
	f.status = regs.status;
	f.fb_width = uint32_t(regs.vi_width);
	f.fb_offset = uint32_t(regs.vi_offset) >> ((regs.status & VI_CONTROL_TYPE_MASK) == VI_CONTROL_TYPE_RGBA8888_BIT ? 2 : 1);
	f.rdram_size = uint32_t(rdram_size);
	f.fetch_bug = (regs.init_y_add < 1024 && scale == 1) ? 1u : 0u; // ref need_fetch_bug_emulation, video_interface.cpp:398-403
	f.scale_log2 = scale_log2;
	f.vram_lo = vram_lo;
	f.vram_hi = vram_hi;
	f.filt_width = std::min<int>(regs.max_x + 2, int(kViStageWidth) * scale);
	f.filt_height = std::min<int>(regs.max_y + 2, int(kViStageHeight) * scale);
	f.stage_stride = kViStageWidth * uint32_t(scale);
	f.layer_stride = f.stage_stride * kViStageHeight * uint32_t(scale);
	(void)divot;

	// Output geometry.
	unsigned width = VI_SCANOUT_WIDTH * unsigned(scale);
	unsigned height = ((regs.is_pal ? VI_V_RES_PAL : VI_V_RES_NTSC) >> int(!serrate)) * unsigned(scale);
	unsigned crop_left, crop_right, crop_top, crop_bottom;
	if (options.crop_rect.enable)
	{
		crop_left = options.crop_rect.left;
		crop_right = options.crop_rect.right;
		crop_top = options.crop_rect.top;
		crop_bottom = options.crop_rect.bottom;
		if (serrate)
		{
			crop_top *= 2;
			crop_bottom *= 2;
		}
	}
	else
	{
		const unsigned crop_y = options.crop_overscan_pixels * (serrate ? 2 : 1);
		const unsigned crop_x = unsigned(roundf(float(crop_y) * (float(width) / float(height))));
		crop_left = crop_right = crop_x;
		crop_top = crop_bottom = crop_y;
	}
	crop_left *= unsigned(scale);
	crop_right *= unsigned(scale);
	crop_top *= unsigned(scale);
	crop_bottom *= unsigned(scale);
	if (crop_left + crop_right < width && crop_top + crop_bottom < height)
	{
		width -= crop_left + crop_right;
		height -= crop_top + crop_bottom;
	}
	else
		fprintf(stderr, "[rdp-b200] Crop of %u x %u is too large for %u x %u.\n", crop_left + crop_right, crop_top + crop_bottom, width, height);

	f.out_width = int(width);
	f.out_height = int(height);
	f.crop_left = int(crop_left);
	f.crop_top = int(crop_top);
	f.y_add = regs.init_y_add;
	f.frame_count = frame_count;
	f.info_y_shift = int(scale_log2);
	if (serrate)
	{
		regs.v_start *= 2;
		regs.v_res *= 2;
		f.serrate_shift = 1;
		f.serrate_mask = 1;
		f.serrate_select = regs.v_current_line == 0 ? 1 : 0;
		f.info_y_shift++;
	}
	// Halving steps of the finished image (ref downscale_stage, video_interface.cpp:1001-1064): at most down to 1x.
	f.downscale_steps = std::min<uint32_t>(std::min(8u, options.downscale_steps), scale_log2);
	// Bob deinterlacing (ref video_interface.cpp:1355-1389): only a serrated frame that was scanned out upscaled as one
	// field (upscale_deinterlacing) and halved gets there; at a remaining scale of 1 the result has twice the rows.
	f.bob_scale = 0;
	f.bob_field = regs.v_current_line == 0 ? 1u : 0u;
	if (f.downscale_steps != 0 && options.upscale_deinterlacing && (regs.status & VI_CONTROL_SERRATE_BIT) != 0)
		f.bob_scale = std::max<uint32_t>(1u, uint32_t(scale) >> f.downscale_steps);
	const unsigned bob_rows = f.bob_scale == 1 ? 2 : 1;
	f.v_start = regs.v_start;

	const int dx = -int(crop_left), dy = -int(crop_top);
	f.draw_x0 = f.draw_y0 = f.draw_x1 = f.draw_y1 = 0;
	if (!degenerate && regs.h_res > 0 && regs.v_res > 0)
		shifted_rect(&f.draw_x0, regs.h_start, regs.v_start, regs.h_res, regs.v_res, dx, dy);

	f.blend_count = 0;
	if (have_prev_scanout && options.blend_previous_frame)
	{
		f.prev_width = int(prev_width);
		f.prev_height = int(prev_height);
		if (degenerate)
		{
			if (regs.h_res > 0 && shifted_rect(f.blend_rect[f.blend_count], regs.h_start, 0, regs.h_res, int(prev_height), dx, dy))
				f.blend_count++;
		}
		else
		{
			if (regs.h_res > 0 && regs.v_start > 0 &&
			    shifted_rect(f.blend_rect[f.blend_count], regs.h_start, 0, regs.h_res, regs.v_start, dx, dy))
				f.blend_count++;
			// The middle part leaves the guard band alone.
			if (regs.h_res_clamp > 0 && regs.v_res > 0 &&
			    shifted_rect(f.blend_rect[f.blend_count], regs.h_start_clamp, regs.v_start, regs.h_res_clamp, regs.v_res, dx, dy))
				f.blend_count++;
			if (regs.h_res > 0 && int(prev_height) > regs.v_start + regs.v_res &&
			    shifted_rect(f.blend_rect[f.blend_count], regs.h_start, regs.v_start + regs.v_res, regs.h_res,
			                 int(prev_height) - (regs.v_start + regs.v_res), dx, dy))
				f.blend_count++;
		}
	}

	if (degenerate)
		f.filt_width = f.filt_height = 0;

	if (rdpb_launch_vi_scanout(device, &f, lines) != 0)
	{
		fprintf(stderr, "[rdp-b200] VI scan-out failed: %s\n", rdpb_last_error(device));
		have_prev_scanout = false;
		return result;
	}

	have_prev_scanout = true;
	prev_width = width;
	prev_height = height;
	prev_result_width = width >> f.downscale_steps;
	prev_result_height = (height >> f.downscale_steps) * bob_rows;
	frame_count++;

	result.device_rgba8 = rdpb_vi_output_device_ptr(device);
	result.width = width >> f.downscale_steps;
	result.height = (height >> f.downscale_steps) * bob_rows;
	return result;
}
}
