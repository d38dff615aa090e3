// rdp_pass.cuh -- per-thread bodies of the render-pass kernels, shared by the CUDA kernels (backend_cuda.cu)
// and the test-only CPU simulator (tests/hostsim).  A pass runs four stages in stream order:
//
//   tmem_update : 2048 threads, one per TMEM word, replay the pass's texture loads and snapshot TMEM after each
//   prim_setup  : one thread per primitive, decode raw command words -> PrimSetup
//   span_bin    : one warp per primitive; lanes own scanlines, edge-walk them into Span records and OR the
//                 primitive's bit into the mask of every 8x8 tile its spans touch (exact, span-derived binning)
//   tiles       : one thread per pixel; load colour/depth once, walk the tile's primitive mask in order,
//                 shade + blend in registers, store once if dirty
//
// Reference dispatch sequence being replaced: parallel-rdp/rdp_renderer.cpp:2223-2285.
#pragma once
#include "rdp_blend.cuh"
#include "rdp_tmem.cuh"

namespace rdpb
{
struct PassBuffers
{
	const uint32_t *words;
	const PrimRecord *records;
	const PassState *states;
	const TmemUpload *uploads;
	PrimSetup *prims;
	Span *spans;
	uint32_t *fine_mask;   // [tile][mask_stride]
	uint32_t *coarse_mask; // [tile][coarse_stride]
	uint32_t *jobs;        // span jobs (span_job_pack), written by the set-up kernel (CUDA backend only)
	uint32_t *job_counter; // [0] span jobs allocated so far in this pass, [1] next tile for the tile kernel's warps
	uint8_t *tmem;           // live TMEM, 4 KiB
	uint8_t *tmem_instances; // snapshot k = TMEM after k uploads of this pass
	uint8_t *rdram;
	uint8_t *hidden;
	const uint8_t *blend_lut;
	unsigned long long *counters; // [0] non-empty tiles processed, [1] shaded (pixel, primitive) pairs
};

// ---- stage 1 -------------------------------------------------------------------------------------
RDP_HD void pass_tmem_word(const PassBuffers &b, uint32_t num_uploads, uint32_t rdram_size, int slot)
{
	uint16_t *live = reinterpret_cast<uint16_t *>(b.tmem);
	uint16_t *inst = reinterpret_cast<uint16_t *>(b.tmem_instances);
	uint32_t value = live[slot];
	inst[slot] = uint16_t(value);

	RdramView rdram;
	rdram.data = b.rdram;
	rdram.mask = rdram_size - 1u;
	for (uint32_t i = 0; i < num_uploads; i++)
	{
		tmem_apply_upload(b.uploads[i], rdram, slot, value);
		inst[size_t(i + 1) * (kTmemBytes / 2) + slot] = uint16_t(value);
	}
	live[slot] = uint16_t(value);
}

// ---- stage 2 -------------------------------------------------------------------------------------
RDP_HD void pass_setup_prim(const PassBuffers &b, uint32_t i)
{
	const PrimRecord rec = b.records[i];
	decode_primitive(b.words + rec.word_offset, b.states[rec.state_id], rec, b.prims[i]);
}

// ---- stage 3 -------------------------------------------------------------------------------------
// Edge-walk scanline y of a primitive.  Returns the inclusive pixel range to bin (lo > hi: nothing).
RDP_HD void pass_span_line(const PassBuffers &b, const FbInfo &fb, const PrimSetup &p, int y, int &bin_lo, int &bin_hi)
{
	bin_lo = 1;
	bin_hi = 0;
	if (y < p.line_lo || y > p.line_hi + 1)
		return;
	Span sp;
	compute_span(p, b.states[p.state_id], y, sp);
	b.spans[p.span_base + uint32_t(y - p.line_lo)] = sp;
	if (y <= p.line_hi && sp.valid && y >= 0 && y < int(fb.height))
	{
		bin_lo = imax(int(sp.start_x), 0);
		bin_hi = imin(int(sp.end_x), int(fb.width) - 1);
	}
}

template <typename OrFn>
RDP_HD void pass_bin_row(const FbInfo &fb, uint32_t prim, int tile_y, int x_lo, int x_hi, OrFn or_fn, uint32_t *fine,
                         uint32_t *coarse)
{
	if (x_lo > x_hi)
		return;
	const uint32_t group = prim >> 5;
	const uint32_t bit = 1u << (prim & 31u);
	const uint32_t cgroup = group >> 5;
	const uint32_t cbit = 1u << (group & 31u);
	for (int tx = x_lo >> 3; tx <= (x_hi >> 3); tx++)
	{
		const uint32_t tile = uint32_t(tile_y) * fb.tiles_x + uint32_t(tx);
		or_fn(&fine[size_t(tile) * fb.mask_stride + group], bit);
		or_fn(&coarse[size_t(tile) * fb.coarse_stride + cgroup], cbit);
	}
}

// ---- stage 4 -------------------------------------------------------------------------------------
template <class S>
RDP_HD void pass_pixel_begin(const PassBuffers &b, const FbInfo &fb, FbView &view, PixelState &px, int x, int y)
{
	view.rdram = b.rdram;
	view.hidden = b.hidden;
	view.blend_lut = b.blend_lut;
	view.fmt = resolve_fb_fmt<S>(fb.fmt);
	view.rdram_size = fb.rdram_size;
	view.alias = fb.color_addr == fb.depth_addr;

	px.r = px.g = px.b = px.a = 0;
	px.depth = 0;
	px.dz = 0;
	px.color_dirty = px.depth_dirty = false;
	px.color_index = 0;
	if (x < int(fb.width) && y < int(fb.height))
	{
		const uint32_t pix = fb.width * uint32_t(y) + uint32_t(x);
		const uint32_t shift = view.fmt == FB_RGBA8888 ? 2u : ((view.fmt == FB_RGBA5551 || view.fmt == FB_IA88) ? 1u : 0u);
		px.color_index = (fb.color_addr >> shift) + pix;
		load_color(view, px, px.color_index);
		load_depth(view, px, (fb.depth_addr >> 1) + pix);
	}
}

RDP_HD void pass_pixel_end(const FbInfo &fb, const FbView &view, PixelState &px, int x, int y)
{
	if (x >= int(fb.width) || y >= int(fb.height))
		return;
	const uint32_t pix = fb.width * uint32_t(y) + uint32_t(x);
	store_color(view, px, px.color_index);
	store_depth(view, px, (fb.depth_addr >> 1) + pix);
}

// Returns the number of primitives that shaded this pixel.
template <class S>
RDP_HD uint32_t pass_pixel_walk(const PassBuffers &b, const FbInfo &fb, const FbView &view, PixelState &px, int x, int y,
                                uint32_t tile)
{
	uint32_t shaded = 0;
	ShadeContext ctx;
	ctx.prims = b.prims;
	ctx.states = b.states;
	ctx.spans = b.spans;
	ctx.tmem = b.tmem_instances;
	ctx.fb_fmt = view.fmt;
	ctx.base_prim_index = fb.base_prim_index;

	const uint32_t *coarse = b.coarse_mask + size_t(tile) * fb.coarse_stride;
	const uint32_t *fine = b.fine_mask + size_t(tile) * fb.mask_stride;
	for (uint32_t cw = 0; cw < fb.coarse_stride; cw++)
	{
		uint32_t cbits = coarse[cw];
		while (cbits)
		{
			const int g = find_lsb(cbits);
			cbits &= cbits - 1u;
			const uint32_t group = cw * 32u + uint32_t(g);
			uint32_t bits = fine[group];
			while (bits)
			{
				const int i = find_lsb(bits);
				bits &= bits - 1u;
				if (process_pixel_primitive<S>(ctx, view, px, x, y, group * 32u + uint32_t(i)))
					shaded++;
			}
		}
	}
	return shaded;
}

// Zero the mask words of one tile after its pixels are done, so the next pass starts from clean masks.
RDP_HD void pass_tile_clear_masks(const PassBuffers &b, const FbInfo &fb, uint32_t tile)
{
	uint32_t *coarse = b.coarse_mask + size_t(tile) * fb.coarse_stride;
	uint32_t *fine = b.fine_mask + size_t(tile) * fb.mask_stride;
	for (uint32_t cw = 0; cw < fb.coarse_stride; cw++)
	{
		uint32_t cbits = coarse[cw];
		if (!cbits)
			continue;
		coarse[cw] = 0;
		while (cbits)
		{
			const int g = find_lsb(cbits);
			cbits &= cbits - 1u;
			fine[cw * 32u + uint32_t(g)] = 0;
		}
	}
}
}
