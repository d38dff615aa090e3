// rdp_pass.cuh -- per-thread bodies of the render-pass kernels, shared by the CUDA kernels (backend_cuda.cu)
// and the test-only CPU simulator (tests/hostsim).  A pass runs four stages in stream order:
//
//   tmem_update : 2048 threads, one per TMEM word, replay the pass's texture loads and snapshot TMEM after each
//   prim_setup  : one thread per primitive, decode raw command words -> PrimSetup
//   span_bin    : one warp per (primitive, 32-line chunk); lanes own scanlines, edge-walk them into Span records and OR
//                 the primitive's bit into the mask of every 8x4 tile its spans touch (exact, span-derived binning)
//   tiles       : one lane per pixel; load colour/depth once, walk the tile's primitive mask in order,
//                 shade + blend in registers, store once if dirty.  pass_pixel_walk is the in-order statement,
//                 pass_pixel_walk_deferred the four-phase one (the CUDA walkers in backend_cuda.cu do the same work
//                 warp-cooperatively)
// plus the PRE / POST synchronisation of the upscaled domain when internal upscaling is on.
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
// With internal upscaling (sl = log2 S) y is a scanline of the upscaled image: a primitive then owns S times as many
// span slots, starting at S * span_base, for lines S * line_lo .. S * (line_hi + 1) + S - 1.
RDP_HD void pass_span_line(const PassBuffers &b, const FbInfo &fb, const PrimSetup &p, int y, int &bin_lo, int &bin_hi, int sl = 0)
{
	bin_lo = 1;
	bin_hi = 0;
	const int S = 1 << sl;
	const int line_lo = int(p.line_lo) << sl, line_hi = (int(p.line_hi) << sl) + (S - 1);
	if (y < line_lo || y > line_hi + S)
		return;
	Span sp;
	compute_span(p, b.states[p.state_id], y, sp, sl);
	b.spans[(p.span_base << sl) + uint32_t(y - line_lo)] = sp;
	if (y <= line_hi && sp.valid && y >= 0 && y < int(fb.height))
	{
		bin_lo = imax(int(int16_t(sp.touch_range & 0xffffu)), 0);
		bin_hi = imin(int(int16_t(sp.touch_range >> 16)), int(fb.width) - 1);
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
// With internal upscaling the pass's RDRAM / hidden RDRAM are arrays of S*S "slices", each a whole image: upscaled
// pixel (x, y) is native pixel (x / S, y / S) of slice (y % S) * S + x % S (memory_interfacing.h:286-302, 56-98).
template <class S>
RDP_HD void pass_pixel_begin(const PassBuffers &b, const FbInfo &fb, FbView &view, PixelState &px, int x, int y, int sl = 0)
{
	const uint32_t slice = (uint32_t(y) & ((1u << sl) - 1u)) * (1u << sl) + (uint32_t(x) & ((1u << sl) - 1u));
	view.rdram = b.rdram + size_t(slice) * fb.rdram_size;
	view.hidden = b.hidden + size_t(slice) * (fb.rdram_size >> 1);
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
		const uint32_t pix = (fb.width >> sl) * (uint32_t(y) >> sl) + (uint32_t(x) >> sl);
		const uint32_t shift = view.fmt == FB_RGBA8888 ? 2u : ((view.fmt == FB_RGBA5551 || view.fmt == FB_IA88) ? 1u : 0u);
		px.color_index = (fb.color_addr >> shift) + pix;
		load_color(view, px, px.color_index);
		load_depth(view, px, (fb.depth_addr >> 1) + pix);
	}
}

RDP_HD void pass_pixel_end(const FbInfo &fb, const FbView &view, PixelState &px, int x, int y, int sl = 0)
{
	if (x >= int(fb.width) || y >= int(fb.height))
		return;
	const uint32_t pix = (fb.width >> sl) * (uint32_t(y) >> sl) + (uint32_t(x) >> sl);
	store_color(view, px, px.color_index);
	store_depth(view, px, (fb.depth_addr >> 1) + pix);
}

// Super-sampled read-back: which native pixel this upscaled pixel stored to.  Returns the bits to OR into *word
// (bit 0 colour, bit 1 depth, at the native pixel's place in its 4x4 block; 0: nothing stored).  The mask words sit
// behind the last RDRAM slice (ref memory_interfacing.h:305-337, finish_tile :340-362).
RDP_HD uint32_t pass_pixel_write_mask(const PassBuffers &b, const FbInfo &fb, const PixelState &px, int x, int y, int sl, uint32_t *&word)
{
	if (!fb.write_mask || x >= int(fb.width) || y >= int(fb.height))
		return 0u;
	const uint32_t bits = (px.color_dirty ? 1u : 0u) | (px.depth_dirty ? 2u : 0u);
	const uint32_t nx = uint32_t(x) >> sl, ny = uint32_t(y) >> sl;
	const uint32_t blocks_x = ((fb.width >> sl) + 3u) >> 2;
	word = reinterpret_cast<uint32_t *>(b.rdram + (size_t(fb.rdram_size) << (2 * sl))) + (ny >> 2) * blocks_x + (nx >> 2);
	return bits << (2u * ((nx & 3u) + 4u * (ny & 3u)));
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
	ctx.scale_log2 = int(fb.scale_log2);

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

// The deferred pipeline (rdp_blend.cuh, defer_*) walked by a single pixel: batches of up to 32 binned primitives go
// through candidate -> depth -> shade -> blend.  This is the scalar statement of what k_tiles_deferred does with
// warp-wide compaction between the phases; the host simulator runs it so the phase split is checked against the
// oracle without a GPU.  Only valid for passes deferrable_pass() accepts.
template <class S>
struct DeferredPixelWalk
{
	const PassBuffers &b;
	const FbInfo &fb;
	const FbView &view;
	PixelState &px;
	int x, y;
	ShadeContext ctx;
	uint32_t prim[32], record[32];
	uint32_t count, shaded;

	RDP_HD void flush()
	{
		uint32_t rx[32], ry[32];
		for (uint32_t i = 0; i < count; i++) // phase C
		{
			const PrimSetup &p = b.prims[prim[i]];
			const Span &sp = b.spans[p.span_base + uint32_t(y - p.line_lo)];
			defer_shade<S>(ctx, x, y, prim[i], p, sp, b.states[p.state_id], record[i], rx[i], ry[i]);
		}
		for (uint32_t i = 0; i < count; i++) // phase D
			defer_blend<S>(view, px, b.states[b.prims[prim[i]].state_id], record[i], rx[i], ry[i]);
		count = 0;
	}

	RDP_HD void add(uint32_t index)
	{
		const PrimSetup &p = b.prims[index];
		if (y < p.line_lo || y > p.line_hi)
			return;
		const Span &sp = b.spans[p.span_base + uint32_t(y - p.line_lo)];
		if (!sp.valid)
			return;
		const PassState &st = b.states[p.state_id];
		const uint32_t cand = defer_candidate<S>(st, p, sp, x); // phase A
		if (!(cand & DC_VALID))
			return;
		shaded++;
		if (cand & DC_FILL) // phase B
		{
			fill_pixel(view, px, st.fill_color);
			count = 0; // the fill overwrites whatever the earlier survivors would have blended
			return;
		}
		const uint32_t rec = defer_depth<S>(view, px, st, p.dz, p.dz_compressed, cand);
		if (!rec)
			return;
		prim[count] = index;
		record[count] = rec;
		count++;
	}
};

template <class S>
RDP_HD uint32_t pass_pixel_walk_deferred(const PassBuffers &b, const FbInfo &fb, const FbView &view, PixelState &px, int x,
                                         int y, uint32_t tile)
{
	DeferredPixelWalk<S> w = { b, fb, view, px, x, y };
	w.ctx.prims = b.prims;
	w.ctx.states = b.states;
	w.ctx.spans = b.spans;
	w.ctx.tmem = b.tmem_instances;
	w.ctx.fb_fmt = view.fmt;
	w.ctx.base_prim_index = fb.base_prim_index;
	w.ctx.scale_log2 = 0; // the deferred pipeline only runs native-resolution passes
	w.count = 0;
	w.shaded = 0;

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
				w.add(group * 32u + uint32_t(i));
			}
			w.flush(); // one fine-mask word = one batch of up to 32 primitives
		}
	}
	return w.shaded;
}

// A pass can take the deferred pipeline when no state in it ties coverage or survival to the shaded colour and
// the colour and depth images are distinct.
RDP_HD bool deferrable_pass(const PassSummary &sum, const FbInfo &fb)
{
	return (sum.rf_or & kRfNotDeferrable) == 0u && fb.color_addr != fb.depth_addr;
}

// ---- internal upscaling: keeping the upscaled domain in step with native RDRAM ------------------------------------
// ref shaders/update_upscaled_domain_pre.comp:95-185 and _post.comp:75-120, called around the native-resolution
// pass (rdp_renderer.cpp:2243-2245, 2268-2277).  `reference` is a native-sized shadow of what the upscaled domain
// believes native RDRAM holds.  PRE, before the native pass renders: wherever the framebuffer or depth range of
// native RDRAM differs from the shadow (the CPU, a texture-like write or another pass touched it), the native value and
// its hidden bits are broadcast to all S*S slices.  POST, after the native pass: the shadow takes the pass's output,
// so that rendered pixels are *not* broadcast over the separately rendered upscaled ones next time.
struct UpscaleSync
{
	uint8_t *rdram, *hidden;       // native
	uint8_t *reference;            // native-sized shadow
	uint8_t *up_rdram, *up_hidden; // S*S slices of rdram_size / (rdram_size >> 1) bytes
	uint32_t rdram_size;
	uint32_t num_slices;
	uint32_t fb_index, depth_index; // first element of the colour image (in pixels of its size) and of the depth image (16-bit)
	uint32_t num_pixels;
	uint32_t size_log2;  // colour pixel size: 0, 1, 2
	uint32_t alias;      // colour and depth share an address
	uint32_t width;      // native framebuffer width (resolve only)
	uint32_t dither;     // resolve only: ordered dither when averaging 5/5/5/1 colours
};

RDP_HD void upscale_sync_word(const UpscaleSync &u, uint32_t index, uint32_t size_log2, bool post)
{
	index &= (u.rdram_size >> size_log2) - 1u;
	const size_t at = size_t(index) << size_log2;
	uint32_t real = 0, ref = 0;
	for (uint32_t k = 0; k < (1u << size_log2); k++)
	{
		real |= uint32_t(u.rdram[at + k]) << (8 * k);
		ref |= uint32_t(u.reference[at + k]) << (8 * k);
	}
	if (real == ref)
		return;
	for (uint32_t k = 0; k < (1u << size_log2); k++)
		u.reference[at + k] = u.rdram[at + k];
	if (post)
		return;
	const size_t hidden_stride = u.rdram_size >> 1;
	for (uint32_t s = 0; s < u.num_slices; s++)
	{
		uint8_t *dst = u.up_rdram + size_t(s) * u.rdram_size;
		uint8_t *hid = u.up_hidden + size_t(s) * hidden_stride;
		for (uint32_t k = 0; k < (1u << size_log2); k++)
			dst[at + k] = u.rdram[at + k];
		if (size_log2 == 0)
		{
			const uint32_t mirrored = index ^ 3u;
			if (mirrored & 1u)
				hid[mirrored >> 1] = u.hidden[mirrored >> 1];
		}
		else if (size_log2 == 1)
			hid[index ^ 1u] = u.hidden[index ^ 1u];
		else
		{
			hid[2u * index] = u.hidden[2u * index];
			hid[2u * index + 1u] = u.hidden[2u * index + 1u];
		}
	}
}

// One thread per framebuffer pixel.
RDP_HD void upscale_sync_pixel(const UpscaleSync &u, uint32_t i, bool post)
{
	if (i >= u.num_pixels)
		return;
	upscale_sync_word(u, u.fb_index + i, u.size_log2, post);
	if (!u.alias)
		upscale_sync_word(u, u.depth_index + i, 1u, post);
}

// Super-sampled read-back (ref shaders/update_upscaled_domain_resolve.comp:94-277, rdp_renderer.cpp:2325-2334): after
// an upscaled-only pass, every native pixel one of whose S*S samples was stored to takes the rounded average of its
// samples' colour channels (16 / 32-bit images; an 8-bit image takes the first sample, as does depth, which cannot be
// filtered), in native RDRAM and in the shadow.  Hidden RDRAM is not resolved.  One thread per native pixel.
RDP_HD void upscale_resolve_pixel(const UpscaleSync &u, uint32_t x, uint32_t y)
{
	if (x >= u.width)
		return;
	const uint32_t i = y * u.width + x;
	if (i >= u.num_pixels)
		return;
	const uint32_t *mask = reinterpret_cast<const uint32_t *>(u.up_rdram + size_t(u.num_slices) * u.rdram_size);
	const uint32_t bits = mask[(y >> 2) * ((u.width + 3u) >> 2) + (x >> 2)] >> (2u * ((x & 3u) + 4u * (y & 3u)));
	const uint32_t n = u.num_slices;
	if (bits & 1u)
	{
		if (u.size_log2 == 0u)
		{
			const uint32_t at = ((u.fb_index + i) & (u.rdram_size - 1u)) ^ 3u;
			const uint8_t v = u.up_rdram[at]; // the reference sums the first sample S*S times
			u.rdram[at] = v;
			u.reference[at] = v;
		}
		else if (u.size_log2 == 1u)
		{
			const uint32_t at = ((u.fb_index + i) & ((u.rdram_size >> 1) - 1u)) ^ 1u;
			uint32_t r = 0, g = 0, b = 0, a = 0;
			for (uint32_t s = 0; s < n; s++)
			{
				const uint32_t w = reinterpret_cast<const uint16_t *>(u.up_rdram + size_t(s) * u.rdram_size)[at];
				r += (w >> 11) & 31u;
				g += (w >> 6) & 31u;
				b += (w >> 1) & 31u;
				a += w & 1u;
			}
			if (u.dither)
			{
				// 4x4 ordered pattern, 3 fractional bits below a 5-bit channel.
				// (rows 0 4 1 5 / 4 0 5 1 / 3 7 2 6 / 7 3 6 2, one nibble per entry)
				const uint32_t d = uint32_t((0x2637627315045140ull >> (4u * ((y & 3u) * 4u + (x & 3u)))) & 15u) * n;
				r = (8u * r + d) / (8u * n);
				g = (8u * g + d) / (8u * n);
				b = (8u * b + d) / (8u * n);
				a = (8u * a + d) / (8u * n);
			}
			else
			{
				r = (r + (n >> 1)) / n;
				g = (g + (n >> 1)) / n;
				b = (b + (n >> 1)) / n;
				a = (a + (n >> 1)) / n;
			}
			const uint16_t v = uint16_t((r << 11) | (g << 6) | (b << 1) | a);
			reinterpret_cast<uint16_t *>(u.rdram)[at] = v;
			reinterpret_cast<uint16_t *>(u.reference)[at] = v;
		}
		else
		{
			const uint32_t at = (u.fb_index + i) & ((u.rdram_size >> 2) - 1u);
			uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
			for (uint32_t s = 0; s < n; s++)
			{
				const uint32_t w = reinterpret_cast<const uint32_t *>(u.up_rdram + size_t(s) * u.rdram_size)[at];
				c3 += w >> 24;
				c2 += (w >> 16) & 255u;
				c1 += (w >> 8) & 255u;
				c0 += w & 255u;
			}
			const uint32_t h = n >> 1;
			const uint32_t v = (((c3 + h) / n) << 24) | (((c2 + h) / n) << 16) | (((c1 + h) / n) << 8) | ((c0 + h) / n);
			reinterpret_cast<uint32_t *>(u.rdram)[at] = v;
			reinterpret_cast<uint32_t *>(u.reference)[at] = v;
		}
	}
	if (!u.alias && (bits & 2u))
	{
		const uint32_t at = ((u.depth_index + i) & ((u.rdram_size >> 1) - 1u)) ^ 1u;
		const uint16_t v = reinterpret_cast<const uint16_t *>(u.up_rdram)[at];
		reinterpret_cast<uint16_t *>(u.rdram)[at] = v;
		reinterpret_cast<uint16_t *>(u.reference)[at] = v;
	}
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
