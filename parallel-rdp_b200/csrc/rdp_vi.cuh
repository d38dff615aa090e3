// rdp_vi.cuh -- VI scan-out, per-pixel bodies shared by the CUDA kernels and the test-only CPU simulator.
//
// Behavioural reference (bit-exact target): parallel-rdp/shaders/extract_vram.comp:61-107, vi_fetch.frag:46-164,
// vi_divot.frag:42-93, vi_scale.frag:49-154, vi_blend_fields.frag:26-32; sequencing video_interface.cpp:405-999.
//
// The reference runs four passes through three intermediate images.  Here the first three collapse into one
// kernel: a "filtered" image F (AA or dither-reconstruction filter, then the divot median) is computed straight
// from RDRAM -- each F pixel re-derives its three horizontal neighbours' fetch results from L1/L2-resident RDRAM
// instead of round-tripping two intermediate images through memory.  A second kernel scales F to the 640-wide
// output, applies gamma / gamma dither and merges the previous field.  F has two layers when the oracle's
// "fetch bug" has to be reproduced (y_add < 0x400).
#pragma once
#include "rdp_setup.cuh"

namespace rdpb
{
enum : uint32_t
{
	// Worst case of the 12-bit scale registers: max_x = (4095 + 640 * 4095) >> 10, max_y = (4095 + 288 * 4095) >> 10.
	kViStageWidth = 2576,
	kViStageHeight = 1168,
	kViMaxLines = 288,
};

enum : uint32_t
{
	VI_STATUS_TYPE_MASK = 3u,
	VI_STATUS_TYPE_RGBA5551 = 2u,
	VI_STATUS_TYPE_RGBA8888 = 3u,
	VI_STATUS_GAMMA_DITHER = 1u << 2,
	VI_STATUS_GAMMA = 1u << 3,
	VI_STATUS_DIVOT = 1u << 4,
	VI_STATUS_SERRATE = 1u << 6,
	VI_STATUS_DITHER_FILTER = 1u << 16,
	VI_STATUS_META_AA = 1u << 17,
	VI_STATUS_META_SCALE = 1u << 18,
};

// Parameters of one scan-out, built by the host (video_interface.cpp) and passed by value to the kernels.
struct ViFrame
{
	uint32_t status;       // VI_STATUS_* with the META bits resolved
	uint32_t fb_offset;    // in pixels of the framebuffer format
	uint32_t fb_width;
	uint32_t rdram_size;
	int32_t filt_width, filt_height; // size of the filtered image F
	uint32_t fetch_bug;    // F has a second layer
	uint32_t stage_stride; // pixels per F row in memory
	uint32_t layer_stride; // pixels between the two layers

	// scale pass
	int32_t out_width, out_height;
	int32_t crop_left, crop_top;
	int32_t v_start, y_add;
	uint32_t frame_count;
	int32_t serrate_shift, serrate_mask, serrate_select, info_y_shift;
	// main draw scissor (already shifted by crop); empty when nothing is drawn
	int32_t draw_x0, draw_y0, draw_x1, draw_y1;
	// previous-frame merge: up to three rectangles (shifted by crop)
	int32_t blend_count;
	int32_t blend_rect[3][4]; // x0, y0, x1, y1
	int32_t prev_width, prev_height; // 0 when there is no previous image

	// Internal upscaling (video_interface.cpp:1299-1317, extract_vram.comp:61-66): the framebuffer is read from the
	// upscaled domain's slices and every coordinate above is in upscaled pixels; the finished image can be halved
	// `downscale_steps` times (video_interface.cpp:1001-1064).
	uint32_t scale_log2;
	uint32_t downscale_steps;
	// Bob deinterlacing of the halved image (video_interface.cpp:1066-1123, 1380-1389): 0 = off, else the scaling
	// factor left after the halving steps (1, 2 or 4); bob_field = this is the field with VCurrentLine == 0.
	uint32_t bob_scale;
	uint32_t bob_field;

	// RDRAM byte range [vram_lo, vram_hi) the filter can read (native addresses; video_interface.cpp:369-396).  The
	// backend runs the scan-out on its own stream and holds back only render passes that write into this range.
	// lo >= hi: unknown, every pass waits.
	uint32_t vram_lo, vram_hi;
};

struct ViBuffers
{
	const uint8_t *rdram;
	const uint8_t *hidden;
	const uint8_t *gamma_lut;
	const ViLine *lines;
	uint32_t *filtered; // F, RGBA8-packed: r | g<<8 | b<<16 | coverage<<24
	uint32_t *out;      // RGBA8, byte order r,g,b,a
	const uint32_t *prev;
};

struct ViColor
{
	int r, g, b, a;
};

RDP_HD uint32_t vi_pack(const ViColor &c)
{
	return uint32_t(c.r & 0xff) | (uint32_t(c.g & 0xff) << 8) | (uint32_t(c.b & 0xff) << 16) | (uint32_t(c.a & 0xff) << 24);
}

RDP_HD ViColor vi_unpack(uint32_t v)
{
	ViColor c;
	c.r = int(v & 0xff);
	c.g = int((v >> 8) & 0xff);
	c.b = int((v >> 16) & 0xff);
	c.a = int(v >> 24);
	return c;
}

// extract_vram.comp: framebuffer pixel at (x, y) -> 8-bit colour + 3-bit coverage.
RDP_HD ViColor vi_read_vram(const ViFrame &f, const ViBuffers &b, int x, int y)
{
	ViColor c;
	const uint32_t type = f.status & VI_STATUS_TYPE_MASK;
	// Upscaled pixel (x, y) is native pixel (x >> sl, y >> sl) of slice (y % S) * S + x % S.
	const int sl = int(f.scale_log2);
	const uint32_t slice = (uint32_t(y) & ((1u << sl) - 1u)) * (1u << sl) + (uint32_t(x) & ((1u << sl) - 1u));
	x >>= sl;
	y >>= sl;
	const uint8_t *rdram = b.rdram + size_t(slice) * f.rdram_size;
	const uint8_t *hidden = b.hidden + size_t(slice) * (f.rdram_size >> 1);
	if (type == VI_STATUS_TYPE_RGBA8888)
	{
		uint32_t idx = uint32_t(y * int(f.fb_width) + x + int(f.fb_offset)) & ((f.rdram_size >> 2) - 1u);
		const uint32_t w = reinterpret_cast<const uint32_t *>(rdram)[idx];
		c.r = int(w >> 24);
		c.g = int((w >> 16) & 0xff);
		c.b = int((w >> 8) & 0xff);
		c.a = int((w >> 5) & 7);
	}
	else if (type == VI_STATUS_TYPE_RGBA5551)
	{
		uint32_t idx = uint32_t(y * int(f.fb_width) + x + int(f.fb_offset)) & ((f.rdram_size >> 1) - 1u);
		const uint32_t w = reinterpret_cast<const uint16_t *>(rdram)[idx ^ 1u];
		c.r = int((w >> 8) & 0xf8u);
		c.g = int((w >> 3) & 0xf8u);
		c.b = int((w << 2) & 0xf8u);
		c.a = int(((w & 1u) << 2) | hidden[idx]);
	}
	else
		c.r = c.g = c.b = c.a = 0;
	if (!(f.status & VI_STATUS_META_AA))
		c.a = 7;
	return c;
}

RDP_HD void vi_track_neighbor(const ViColor &n, int *lo, int *hi, int *lo2, int *hi2)
{
	if (n.a != 7)
		return;
	const int v[3] = { n.r, n.g, n.b };
	for (int i = 0; i < 3; i++)
	{
		lo2[i] = imin(lo2[i], imax(v[i], lo[i]));
		hi2[i] = imax(hi2[i], imin(v[i], hi[i]));
		lo[i] = imin(v[i], lo[i]);
		hi[i] = imax(v[i], hi[i]);
	}
}

RDP_HD int vi_clamp1(int v) { return iclamp(v, -1, 1); }

// vi_fetch.frag: AA filter (partial coverage) or dither reconstruction (full coverage) at framebuffer pixel (x, y).
// `bug` selects the variant the oracle produces when it re-reads a line ("fetch bug").
//
// The two halves of vi_fetch.frag, split so that the CUDA filter kernel can run the (rare, expensive) partial-coverage
// half over a compacted list of pixels instead of making every warp that contains one such pixel walk through it.
//
// Partial coverage (mid.a != 7): the AA filter.
template <class Reader>
RDP_HD ViColor vi_fetch_partial(const ViFrame &f, const Reader &read, int x, int y, bool bug, const ViColor &mid)
{
	ViColor out = mid;
	int lo[3] = { mid.r, mid.g, mid.b }, hi[3] = { mid.r, mid.g, mid.b };
	int lo2[3] = { mid.r, mid.g, mid.b }, hi2[3] = { mid.r, mid.g, mid.b };
	const ViColor to_left = read(x - 2, y);
	const ViColor to_right = read(x + 2, y);
	vi_track_neighbor(read(x - 1, y - 1), lo, hi, lo2, hi2);
	vi_track_neighbor(read(x + 1, y - 1), lo, hi, lo2, hi2);
	vi_track_neighbor(to_left, lo, hi, lo2, hi2);
	vi_track_neighbor(to_right, lo, hi, lo2, hi2);
	if (!bug)
	{
		vi_track_neighbor(read(x - 1, y + 1), lo, hi, lo2, hi2);
		vi_track_neighbor(read(x + 1, y + 1), lo, hi, lo2, hi2);
	}
	else
	{
		// The lower row is not read; left/right are seen twice instead.
		vi_track_neighbor(to_left, lo, hi, lo2, hi2);
		vi_track_neighbor(to_right, lo, hi, lo2, hi2);
	}
	const int m[3] = { mid.r, mid.g, mid.b };
	int res[3];
	for (int i = 0; i < 3; i++)
	{
		if (f.fetch_bug)
		{
			// Only compiled in when the bug layer exists (FETCH_BUG define of the reference shader).
			if (m[i] == lo[i]) lo2[i] = lo[i];
			if (m[i] == hi[i]) hi2[i] = hi[i];
		}
		const uint32_t offset = uint32_t(lo2[i] + hi2[i] - (m[i] << 1));
		const uint32_t coeff = uint32_t(7 - mid.a);
		res[i] = int((uint32_t(m[i]) + ((offset * coeff + 4u) >> 3)) & 0xffu);
	}
	out.r = res[0];
	out.g = res[1];
	out.b = res[2];
	return out;
}

// Full coverage: dither reconstruction (when enabled), else the pixel as it is.
template <class Reader>
RDP_HD ViColor vi_fetch_full(const ViFrame &f, const Reader &read, int x, int y, bool bug, const ViColor &mid)
{
	ViColor out = mid;
	if (f.status & VI_STATUS_DITHER_FILTER)
	{
		const int base[3] = { mid.r >> 3, mid.g >> 3, mid.b >> 3 };
		int acc[3] = { 0, 0, 0 };
		for (int dy = -1; dy <= 0; dy++)
		{
			for (int dx = -1; dx <= 1; dx++)
			{
				const ViColor c = read(x + dx, y + dy);
				acc[0] += vi_clamp1((c.r >> 3) - base[0]);
				acc[1] += vi_clamp1((c.g >> 3) - base[1]);
				acc[2] += vi_clamp1((c.b >> 3) - base[2]);
			}
		}
		if (!bug)
		{
			for (int dx = -1; dx <= 1; dx++)
			{
				const ViColor c = read(x + dx, y + 1);
				acc[0] += vi_clamp1((c.r >> 3) - base[0]);
				acc[1] += vi_clamp1((c.g >> 3) - base[1]);
				acc[2] += vi_clamp1((c.b >> 3) - base[2]);
			}
		}
		else
		{
			for (int dx = -1; dx <= 1; dx += 2)
			{
				const ViColor c = read(x + dx, y);
				acc[0] += vi_clamp1((c.r >> 3) - base[0]);
				acc[1] += vi_clamp1((c.g >> 3) - base[1]);
				acc[2] += vi_clamp1((c.b >> 3) - base[2]);
			}
		}
		// The reference shader masks the centre to 5 bits here (vi_fetch.frag:144); that is a no-op for 16-bit
		// framebuffers but differs from the oracle for 32-bit ones (angrylion vi/restore.c:98-140 keeps all 8
		// bits).  We follow the oracle.
		out.r = mid.r + acc[0];
		out.g = mid.g + acc[1];
		out.b = mid.b + acc[2];
	
	}
	return out;
}

// `read(x, y)` yields the decoded framebuffer pixel: straight from RDRAM (vi_fetch_pixel below), or from a block's
// shared-memory copy of its neighbourhood in the CUDA filter kernel, where every pixel is then decoded once
// instead of once per consumer.
template <class Reader>
RDP_HD ViColor vi_fetch_core(const ViFrame &f, const Reader &read, int x, int y, bool bug)
{
	const ViColor mid = read(x, y);
	return mid.a != 7 ? vi_fetch_partial(f, read, x, y, bug, mid) : vi_fetch_full(f, read, x, y, bug, mid);
}

struct ViVramReader
{
	const ViFrame &f;
	const ViBuffers &b;
	RDP_HD ViColor operator()(int x, int y) const { return vi_read_vram(f, b, x, y); }
};

RDP_HD ViColor vi_fetch_pixel(const ViFrame &f, const ViBuffers &b, int x, int y, bool bug)
{
	const ViVramReader read = { f, b };
	return vi_fetch_core(f, read, x, y, bug);
}

RDP_HD int vi_median3(int l, int c, int r)
{
	return imax(imin(l, c), imin(imax(l, c), r));
}

// vi_divot.frag on top of vi_fetch: filtered pixel for framebuffer position (x, y).
RDP_HD ViColor vi_filtered_pixel(const ViFrame &f, const ViBuffers &b, int x, int y, bool bug)
{
	const ViColor mid = vi_fetch_pixel(f, b, x, y, bug);
	if (!(f.status & VI_STATUS_DIVOT))
		return mid;
	const ViColor left = vi_fetch_pixel(f, b, x - 1, y, bug);
	const ViColor right = vi_fetch_pixel(f, b, x + 1, y, bug);
	if ((left.a & mid.a & right.a) == 7)
		return mid;
	ViColor out;
	out.r = vi_median3(left.r, mid.r, right.r);
	out.g = vi_median3(left.g, mid.g, right.g);
	out.b = vi_median3(left.b, mid.b, right.b);
	out.a = mid.a;
	return out;
}

// Kernel 1 body: one thread per F pixel (and layer).
RDP_HD void vi_filter_thread(const ViFrame &f, const ViBuffers &b, int x, int y, int layer)
{
	const ViColor c = vi_filtered_pixel(f, b, x, y, layer != 0);
	b.filtered[size_t(layer) * f.layer_stride + size_t(y) * f.stage_stride + size_t(x)] = vi_pack(c);
}

RDP_HD ViColor vi_load_filtered(const ViFrame &f, const ViBuffers &b, int x, int y, int layer)
{
	if (x < 0 || y < 0 || x >= f.filt_width || y >= f.filt_height)
	{
		ViColor z;
		z.r = z.g = z.b = z.a = 0;
		return z;
	}
	return vi_unpack(b.filtered[size_t(layer) * f.layer_stride + size_t(y) * f.stage_stride + size_t(x)]);
}

RDP_HD int vi_lerp(int a, int b, int l)
{
	return int((uint32_t(a) + ((uint32_t(b - a) * uint32_t(l) + 16u) >> 5)) & 0xffu);
}

RDP_HD uint32_t vi_noise(uint32_t x, uint32_t y, uint32_t frame)
{
	const uint32_t P = 1103515245u;
	uint32_t a = x, b = y, c = frame;
	for (int i = 0; i < 3; i++)
	{
		const uint32_t na = ((a >> 8) ^ b) * P, nb = ((b >> 8) ^ c) * P, nc = ((c >> 8) ^ a) * P;
		a = na;
		b = nb;
		c = nc;
	}
	return (a >> 16) & 0xffffu;
}

RDP_HD bool vi_in_rect(const int32_t *r, int x, int y) { return x >= r[0] && y >= r[1] && x < r[2] && y < r[3]; }

// Kernel 2 body: one thread per output pixel.  vi_scale.frag + vi_blend_fields.frag.
RDP_HD void vi_scale_thread(const ViFrame &f, const ViBuffers &b, int ox, int oy)
{
	uint32_t result = 0; // cleared: rgb = 0, alpha = 0
	bool written = false;

	if (vi_in_rect(&f.draw_x0, ox, oy))
	{
		int cx = ox + f.crop_left, cy = oy + f.crop_top;
		const ViLine &line = b.lines[imin(cy >> f.info_y_shift, int(kViMaxLines) - 1)];
		cy -= f.v_start;
		if (!(cx < line.h_start_clamp || cx >= line.h_end_clamp || (cy & f.serrate_mask) != f.serrate_select))
		{
			cx -= line.h_start;
			cy >>= f.serrate_shift;
			uint32_t noise = 0;
			if (f.status & VI_STATUS_GAMMA_DITHER)
				noise = vi_noise(uint32_t(cx), uint32_t(cy), f.frame_count);

			const int x = cx * line.x_add + line.x_start;
			const int y = (cy - line.y_base) * line.y_add + line.y_start;
			const int bx = x >> 10, by = y >> 10;
			ViColor c = vi_load_filtered(f, b, bx, by, 0);

			int bug_layer = 0;
			if (f.fetch_bug)
			{
				const int prev_y = (y - f.y_add) >> 10;
				const int next_y = (y + f.y_add) >> 10;
				if (cy != 0 && by == prev_y && by != next_y)
					bug_layer = 1;
			}

			if (f.status & VI_STATUS_META_SCALE)
			{
				const int xf = (x >> 5) & 31, yf = (y >> 5) & 31;
				const ViColor c10 = vi_load_filtered(f, b, bx + 1, by, 0);
				const ViColor c01 = vi_load_filtered(f, b, bx, by + 1, bug_layer);
				const ViColor c11 = vi_load_filtered(f, b, bx + 1, by + 1, bug_layer);
				const int r0 = vi_lerp(c.r, c01.r, yf), g0 = vi_lerp(c.g, c01.g, yf), b0 = vi_lerp(c.b, c01.b, yf);
				const int r1 = vi_lerp(c10.r, c11.r, yf), g1 = vi_lerp(c10.g, c11.g, yf), b1 = vi_lerp(c10.b, c11.b, yf);
				c.r = vi_lerp(r0, r1, xf);
				c.g = vi_lerp(g0, g1, xf);
				c.b = vi_lerp(b0, b1, xf);
			}

			if (f.status & VI_STATUS_GAMMA)
			{
				if (f.status & VI_STATUS_GAMMA_DITHER)
				{
					const uint32_t dr = noise & 0x3f, dg = (noise >> 6) & 0x3f, db = ((noise >> 9) & 0x38) | (noise & 7u);
					c.r = b.gamma_lut[(uint32_t(c.r) << 6) + dr + 256u];
					c.g = b.gamma_lut[(uint32_t(c.g) << 6) + dg + 256u];
					c.b = b.gamma_lut[(uint32_t(c.b) << 6) + db + 256u];
				}
				else
				{
					c.r = b.gamma_lut[c.r];
					c.g = b.gamma_lut[c.g];
					c.b = b.gamma_lut[c.b];
				}
			}
			else if (f.status & VI_STATUS_GAMMA_DITHER)
			{
				c.r = imin(c.r + int(noise & 1u), 0xff);
				c.g = imin(c.g + int((noise >> 1) & 1u), 0xff);
				c.b = imin(c.b + int((noise >> 2) & 1u), 0xff);
			}
			c.a = 0xff;
			result = vi_pack(c);
			written = true;
		}
	}

	if (!written && f.prev_width)
	{
		// Previous frame persists for exactly one frame where this frame did not write.
		bool in_blend = false;
		for (int i = 0; i < f.blend_count; i++)
			in_blend = in_blend || vi_in_rect(f.blend_rect[i], ox, oy);
		if (in_blend && ox < f.prev_width && oy < f.prev_height)
		{
			const uint32_t p = b.prev[size_t(oy) * size_t(f.prev_width) + size_t(ox)];
			if (p >> 24)
				result = p & 0x00ffffffu;
		}
	}
	b.out[size_t(oy) * size_t(f.out_width) + size_t(ox)] = result;
}

// One halving step of the finished upscaled image: each output pixel is the rounded mean of a 2x2 block (the
// reference blits with a linear filter at exactly half size, video_interface.cpp:1054-1057).
RDP_HD void vi_downscale_thread(const uint32_t *src, int src_width, uint32_t *dst, int dst_width, int x, int y)
{
	const uint32_t p00 = src[size_t(2 * y) * size_t(src_width) + size_t(2 * x)];
	const uint32_t p10 = src[size_t(2 * y) * size_t(src_width) + size_t(2 * x + 1)];
	const uint32_t p01 = src[size_t(2 * y + 1) * size_t(src_width) + size_t(2 * x)];
	const uint32_t p11 = src[size_t(2 * y + 1) * size_t(src_width) + size_t(2 * x + 1)];
	uint32_t out = 0;
	for (int sh = 0; sh < 32; sh += 8)
		out |= ((((p00 >> sh) & 0xffu) + ((p10 >> sh) & 0xffu) + ((p01 >> sh) & 0xffu) + ((p11 >> sh) & 0xffu) + 2u) >> 2) << sh;
	dst[size_t(y) * size_t(dst_width) + size_t(x)] = out;
}

// Bob deinterlacing (ref upscale_deinterlace, video_interface.cpp:1066-1123 with vi_deinterlace.vert/.frag): the one
// field of a serrated frame, scaled and halved, is resampled with a clamped linear filter at a vertical offset of
// -+ scale / 4 source rows; at scale 1 the output also has twice the rows.  In texel space the sample position of output
// row y is y / 2 - 1/2 or y / 2 (scale 1), y -+ scale / 4 (scale 2, 4): always a whole or a half row, so the filter is
// either one source row or the mean of two neighbours (ties round up; the reference leaves that to the sampler).
RDP_HD void vi_bob_thread(const uint32_t *src, int width, int src_height, uint32_t *dst, uint32_t scale, bool field, int x, int y)
{
	// twice the sample row
	const int ty2 = scale == 1u ? (field ? y - 1 : y) : (field ? 2 * y - int(scale >> 1) : 2 * y + int(scale >> 1));
	int r0 = ty2 >> 1;
	int r1 = r0 + (ty2 & 1);
	r0 = r0 < 0 ? 0 : (r0 >= src_height ? src_height - 1 : r0);
	r1 = r1 < 0 ? 0 : (r1 >= src_height ? src_height - 1 : r1);
	const uint32_t a = src[size_t(r0) * size_t(width) + size_t(x)];
	const uint32_t b = src[size_t(r1) * size_t(width) + size_t(x)];
	uint32_t out = 0;
	for (int sh = 0; sh < 32; sh += 8)
		out |= ((((a >> sh) & 0xffu) + ((b >> sh) & 0xffu) + 1u) >> 1) << sh;
	dst[size_t(y) * size_t(width) + size_t(x)] = out;
}

void build_blender_lut(uint8_t *lut);
void build_gamma_lut(uint8_t *lut);
}
