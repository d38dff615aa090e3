// rdp_renderer.cpp -- see rdp_renderer.hpp.
//
// Reference behaviour this file re-implements (not copies): parallel-rdp/rdp_renderer.cpp
//   state deduction      :1063-1160 (texel/LOD/noise usage), :1294-1400
//   flush rules          :679-695 (image changes), :1564-1590 (capacity), :3051-3097 (texture-from-framebuffer)
//   pass submission      :2223-2285, :2909-2986
//   texture load split   :3099-3486
#include "rdp_renderer.hpp"
#include "rdp_setup.cuh"
#include <algorithm>
#include <stdio.h>
#include <string.h>

using namespace rdpb;

namespace RDP
{
void Renderer::set_quirks(uint32_t quirk_flags)
{
	if (regs.quirks != quirk_flags)
	{
		regs.quirks = quirk_flags;
		state_dirty = true;
	}
}

bool Renderer::set_upscaling(uint32_t log2)
{
	if (log2 == 0)
		return true;
	if (log2 > 3 || rdpb_upscaling_init(device, log2) != 0)
	{
		fprintf(stderr, "[rdp-b200] Cannot set up %ux internal upscaling; rendering at native resolution only.\n", 1u << log2);
		return false;
	}
	upscale_log2 = log2;
	return true;
}

bool Renderer::set_super_sampled_readback(bool enable, bool dither)
{
	// ref rdp_renderer.cpp:60-61: super-sampling a 1x render is refused.
	if (enable && upscale_log2 == 0)
		return false;
	super_sampled_readback = enable;
	super_sampled_dither = enable && dither;
	return true;
}

bool Renderer::init(rdpb_device *device_, size_t rdram_size_)
{
	device = device_;
	rdram_size = rdram_size_;
	memset(&regs, 0, sizeof(regs));
	state_dirty = true;
	return device != nullptr;
}

// ---------------------------------------------------------------------------------------------
// Register decode helpers (SetOtherModes / SetCombine): hardware bit positions as in rdp_device.cpp:474-540, :661-684
// ---------------------------------------------------------------------------------------------
void Renderer::set_other_modes(uint32_t w0, uint32_t w1)
{
	other_modes[0] = w0;
	other_modes[1] = w1;
	state_dirty = true;
}

void Renderer::set_combine(uint32_t w0, uint32_t w1)
{
	// cycle 0
	regs.comb_rgb[0][0] = (w0 >> 20) & 0xf;
	regs.comb_rgb[0][1] = (w1 >> 28) & 0xf;
	regs.comb_rgb[0][2] = (w0 >> 15) & 0x1f;
	regs.comb_rgb[0][3] = (w1 >> 15) & 0x7;
	regs.comb_alpha[0][0] = (w0 >> 12) & 0x7;
	regs.comb_alpha[0][1] = (w1 >> 12) & 0x7;
	regs.comb_alpha[0][2] = (w0 >> 9) & 0x7;
	regs.comb_alpha[0][3] = (w1 >> 9) & 0x7;
	// cycle 1
	regs.comb_rgb[1][0] = (w0 >> 5) & 0xf;
	regs.comb_rgb[1][1] = (w1 >> 24) & 0xf;
	regs.comb_rgb[1][2] = (w0 >> 0) & 0x1f;
	regs.comb_rgb[1][3] = (w1 >> 6) & 0x7;
	regs.comb_alpha[1][0] = (w1 >> 21) & 0x7;
	regs.comb_alpha[1][1] = (w1 >> 3) & 0x7;
	regs.comb_alpha[1][2] = (w1 >> 18) & 0x7;
	regs.comb_alpha[1][3] = (w1 >> 0) & 0x7;
	state_dirty = true;
}

static bool cycle_reads_texel(const PassState &st, int cycle, int which)
{
	const uint8_t *rgb = st.comb_rgb[cycle];
	const uint8_t *a = st.comb_alpha[cycle];
	const int sel = which + 1;      // TEXEL0 = 1, TEXEL1 = 2 in every selector table
	const int sel_alpha = which + 8; // TEXEL0_ALPHA = 8, TEXEL1_ALPHA = 9 in the RGB multiplier table
	return rgb[0] == sel || rgb[1] == sel || rgb[2] == sel || rgb[3] == sel || rgb[2] == sel_alpha || a[0] == sel ||
	       a[1] == sel || a[2] == sel || a[3] == sel;
}

static bool cycle_reads_lod_frac(const PassState &st, int cycle)
{
	return st.comb_rgb[cycle][2] == 13 || st.comb_alpha[cycle][2] == 0;
}

// Fold other-modes, scissor field bits and combiner usage into the flags the kernels consume.
void Renderer::finalize_state(PassState &st) const
{
	const uint32_t w0 = other_modes[0], w1 = other_modes[1];
	uint32_t rf = st.raster_flags & (RF_INTERLACE | RF_KEEP_ODD); // owned by SetScissor
	uint32_t bf = 0;

	if (w0 & (1u << 19)) rf |= RF_PERSPECTIVE;
	if (w0 & (1u << 18)) rf |= RF_DETAIL;
	if (w0 & (1u << 17)) rf |= RF_SHARPEN;
	if (w0 & (1u << 16)) rf |= RF_TEX_LOD;
	if (w0 & (1u << 15)) rf |= RF_TLUT;
	if (w0 & (1u << 14)) rf |= RF_TLUT_IA;
	if (w0 & (1u << 13)) rf |= RF_SAMPLE_QUAD;
	if (w0 & (1u << 12)) rf |= RF_MID_TEXEL;
	if (w0 & (1u << 11)) rf |= RF_BILERP0;
	if (w0 & (1u << 10)) rf |= RF_BILERP1;
	if (w0 & (1u << 9)) rf |= RF_CONVERT_ONE;
	if (w1 & (1u << 14)) bf |= BF_FORCE_BLEND;
	if (w1 & (1u << 13)) rf |= RF_ALPHA_CVG_SELECT;
	if (w1 & (1u << 12)) rf |= RF_CVG_TIMES_ALPHA;
	if (w1 & (1u << 7)) bf |= BF_COLOR_ON_CVG;
	if (w1 & (1u << 6)) bf |= BF_IMAGE_READ;
	if (w1 & (1u << 5)) bf |= BF_Z_UPDATE;
	if (w1 & (1u << 4)) bf |= BF_Z_COMPARE;
	if (w1 & (1u << 3))
	{
		rf |= RF_AA;
		bf |= BF_AA;
	}
	if (w1 & (1u << 2)) rf |= RF_PRIM_DEPTH;
	if (w1 & (1u << 1)) rf |= RF_ALPHA_TEST_DITHER;
	if (w1 & (1u << 0)) rf |= RF_ALPHA_TEST;

	st.dither_rgb = (w0 >> 6) & 3;
	st.dither_alpha = (w0 >> 4) & 3;
	if (st.dither_rgb != 3)
		bf |= BF_DITHER;
	st.cvg_mode = (w1 >> 8) & 3;
	st.z_mode = (w1 >> 10) & 3;

	switch ((w0 >> 20) & 3)
	{
	case 1: rf |= RF_TWO_CYCLE; bf |= BF_TWO_CYCLE; break;
	case 2: rf |= RF_COPY; break;
	case 3: rf |= RF_FILL; break;
	default: break;
	}

	st.blend_modes[0][0] = (w1 >> 30) & 3;
	st.blend_modes[1][0] = (w1 >> 28) & 3;
	st.blend_modes[0][1] = (w1 >> 26) & 3;
	st.blend_modes[1][1] = (w1 >> 24) & 3;
	st.blend_modes[0][2] = (w1 >> 22) & 3;
	st.blend_modes[1][2] = (w1 >> 20) & 3;
	st.blend_modes[0][3] = (w1 >> 18) & 3;
	st.blend_modes[1][3] = (w1 >> 16) & 3;

	// Which texels / LOD the combiner can observe.
	if (!(rf & (RF_FILL | RF_COPY)))
	{
		bool t0, t1, pipelined_t1, lod;
		if (rf & RF_TWO_CYCLE)
		{
			// The second cycle sees TEXEL0/TEXEL1 swapped.
			t0 = cycle_reads_texel(st, 0, 0) || cycle_reads_texel(st, 1, 1);
			t1 = cycle_reads_texel(st, 0, 1) || cycle_reads_texel(st, 1, 0);
			pipelined_t1 = false;
			lod = cycle_reads_lod_frac(st, 0) || cycle_reads_lod_frac(st, 1);
		}
		else
		{
			t0 = cycle_reads_texel(st, 1, 0);
			t1 = false;
			pipelined_t1 = cycle_reads_texel(st, 1, 1);
			lod = false;
		}
		if (t1 && (rf & RF_CONVERT_ONE))
			t0 = true;
		if (t0) rf |= RF_USES_TEXEL0;
		if (t1) rf |= RF_USES_TEXEL1;
		if (pipelined_t1) rf |= RF_USES_PIPELINED_TEXEL1;
		if (lod || (rf & RF_TEX_LOD)) rf |= RF_USES_LOD;
	}

	// Noise usage.
	if (st.dither_alpha == 2 || st.dither_rgb == 2)
		rf |= RF_NOISE;
	else if (!(rf & (RF_FILL | RF_COPY)))
	{
		const bool two = (rf & RF_TWO_CYCLE) != 0;
		const bool n0 = two && st.comb_rgb[0][0] == 7;
		const bool n1 = st.comb_rgb[1][0] == 7;
		if (n0 || n1)
			rf |= RF_NOISE;
		if (n0 && n1)
			rf |= RF_NOISE_DUAL;
		if ((rf & RF_ALPHA_TEST) && (rf & RF_ALPHA_TEST_DITHER))
			rf |= RF_NOISE;
	}

	st.raster_flags = rf;
	st.blend_flags = bf;
}

void Renderer::ensure_pass()
{
	if (pass_open)
		return;
	{
		// May block until the staging slot's previous uploads (four passes ago) have left host memory: time spent here
		// is back-pressure from the GPU, not command-thread work, and is reported apart (get_backpressure_ns).
		const auto t0 = std::chrono::steady_clock::now();
		rdpb_pass_begin(device, &pass);
		backpressure_ns += uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
	}
	pass_open = true;
	num_words = num_prims = num_states = num_uploads = num_spans = num_jobs = 0;
	staged_words = 0;
	segments.clear();
	pass_reads_tmem = false;
	reset_summary();
	state_dirty = true;
	memset(state_table, 0, sizeof(state_table));
	current_state = 0;
}

void Renderer::reset_summary()
{
	summary = {};
	summary.rf_and = ~0u;
	summary.bf_and = ~0u;
	summary.modes_uniform = 1;
}

void Renderer::snapshot_state()
{
	// Build the snapshot in the next free slot; keep the slot only if the pass has no identical one yet.
	PassState *states = static_cast<PassState *>(pass.states);
	PassState &dst = states[num_states];
	dst = regs;
	finalize_state(dst);
	state_dirty = false;

	static_assert(sizeof(PassState) % 8 == 0, "hashed as 64-bit words");
	uint64_t h = 0xcbf29ce484222325ull;
	{
		uint64_t words[sizeof(PassState) / 8];
		memcpy(words, &dst, sizeof(PassState));
		for (uint64_t w : words)
			h = (h ^ w) * 0x100000001b3ull + (h >> 29);
	}
	uint32_t slot = uint32_t(h >> 20) & (StateTableSize - 1);
	while (state_table[slot])
	{
		const uint32_t idx = state_table[slot] - 1u;
		if (state_hash[idx] == h && memcmp(&states[idx], &dst, sizeof(PassState)) == 0)
		{
			current_state = idx;
			const bool found_tlut = (states[idx].raster_flags & RF_TLUT) != 0;
			for (int i = 0; i < 8; i++)
				state_tile_kind_bits[i] = 1u << resolve_texel_kind(states[idx].tiles[i].fmt, states[idx].tiles[i].size, found_tlut);
			return; // already counted in the summary
		}
		slot = (slot + 1) & (StateTableSize - 1);
	}
	state_table[slot] = uint16_t(num_states + 1u);
	state_hash[num_states] = h;
	current_state = num_states++;

	// Fill-mode states are handled by every tile-kernel instantiation; only the others constrain the choice.
	if (dst.raster_flags & RF_FILL)
		return;
	const uint64_t comb_rgb = pack_comb_rgb(dst.comb_rgb);
	const uint32_t comb_alpha = pack_comb_alpha(dst.comb_alpha);
	const uint32_t blend_misc = pack_blend_misc(dst.blend_modes, dst.dither_rgb, dst.dither_alpha, dst.cvg_mode, dst.z_mode);
	if (summary.num_shaded_states == 0)
	{
		summary.comb_rgb = comb_rgb;
		summary.comb_alpha = comb_alpha;
		summary.blend_misc = blend_misc;
	}
	else if (summary.comb_rgb != comb_rgb || summary.comb_alpha != comb_alpha || summary.blend_misc != blend_misc)
		summary.modes_uniform = 0;
	summary.num_shaded_states++;
	summary.rf_or |= dst.raster_flags;
	summary.rf_and &= dst.raster_flags;
	summary.bf_or |= dst.blend_flags;
	summary.bf_and &= dst.blend_flags;

	const bool tlut = (dst.raster_flags & RF_TLUT) != 0;
	for (int i = 0; i < 8; i++)
		state_tile_kind_bits[i] = 1u << resolve_texel_kind(dst.tiles[i].fmt, dst.tiles[i].size, tlut);
}

void Renderer::validate_draw_state() const
{
	const uint32_t cycle = (other_modes[0] >> 20) & 3;
	const uint32_t w1 = other_modes[1];
	if (cycle == 3)
	{
		if (fb.fmt == FBFormat::I4)
			validation->report_rdp_crash(ValidationError::Fill4bpp, "Attempted to use Fill mode on 4bpp surface.");
		if (w1 & (1u << 4))
			validation->report_rdp_crash(ValidationError::FillDepthTest, "Attempted to use Fill mode with depth test.");
		if (w1 & (1u << 6))
			validation->report_rdp_crash(ValidationError::FillImageReadEnable, "Attempted to use Fill mode with image read enable.");
		if ((w1 & (1u << 5)) && !(w1 & (1u << 2)))
			validation->report_rdp_crash(ValidationError::FillDepthWrite, "Attempted to use Fill mode with depth write enabled.");
	}
	else if (cycle == 2 && fb.fmt == FBFormat::RGBA8888)
		validation->report_rdp_crash(ValidationError::Copy32bpp, "Attempted to use Copy mode on 32bpp surface.");
}

void Renderer::set_color_framebuffer(uint32_t addr, uint32_t width, FBFormat fmt)
{
	if (fb.addr != addr || fb.width != width || fb.fmt != fmt)
		flush_queues();
	fb.addr = addr;
	fb.width = width;
	fb.fmt = fmt;
}

void Renderer::set_depth_framebuffer(uint32_t addr)
{
	if (fb.depth_addr != addr)
		flush_queues();
	fb.depth_addr = addr;
}

// Decide where a primitive's words live in the device word buffer; returns that offset.
uint32_t Renderer::place_words(const uint32_t *words, unsigned count)
{
	enum : uint32_t { kMaxGapWords = 1024 };
	const bool borrowed = words >= borrowed_base && words + count <= borrowed_end;
	if (borrowed)
	{
		if (!segments.empty())
		{
			WordSegment &seg = segments.back();
			const uint32_t *seg_end = seg.src + seg.count;
			if (seg.borrowed && seg.src >= borrowed_base && words >= seg_end && size_t(words - seg_end) <= kMaxGapWords)
			{
				const uint32_t offset = seg.dst + uint32_t(words - seg.src);
				seg.count = uint32_t(words - seg.src) + count;
				num_words = seg.dst + seg.count;
				return offset;
			}
		}
		segments.push_back({ words, count, num_words, true });
		const uint32_t offset = num_words;
		num_words += count;
		return offset;
	}

	uint32_t *dst = pass.words + staged_words;
	memcpy(dst, words, count * sizeof(uint32_t));
	if (!segments.empty() && !segments.back().borrowed && segments.back().src + segments.back().count == dst)
		segments.back().count += count;
	else
		segments.push_back({ dst, count, num_words, false });
	staged_words += count;
	const uint32_t offset = num_words;
	num_words += count;
	return offset;
}

void Renderer::draw_primitive(const uint32_t *words, unsigned count)
{
	if (validation)
		validate_draw_state();

	ensure_pass();
	if (state_dirty)
	{
		if (num_states == kMaxStatesPerPass)
		{
			flush_queues();
			ensure_pass();
		}
		snapshot_state();
	}

	const PassState &st = static_cast<const PassState *>(pass.states)[current_state];

	int line_lo, line_hi;
	primitive_line_range(words, st.raster_flags, st.sc_ylo, st.sc_yhi, line_lo, line_hi);

	PrimRecord &rec = static_cast<PrimRecord *>(pass.records)[num_prims];
	rec.word_offset = place_words(words, count);
	rec.span_base = num_spans;
	rec.state_id = uint16_t(current_state);
	rec.tmem_instance = uint16_t(num_uploads);
	rec.line_lo = int16_t(line_lo);
	rec.line_hi = int16_t(line_hi);

	num_prims++;

	if (line_hi >= line_lo)
	{
		num_spans += uint32_t(line_hi - line_lo + 2);
		num_jobs += span_job_count(line_lo, line_hi);
		fb.deduced_height = std::max(fb.deduced_height, uint32_t(line_hi + 1));
	}

	if (!(st.raster_flags & RF_FILL))
	{
		pass_reads_tmem = true;
		// Texel kinds this primitive can sample: its own tile (+1 for the second texel), or any tile under LOD.
		const uint32_t rf = st.raster_flags;
		if (rf & (RF_USES_TEXEL0 | RF_USES_TEXEL1 | RF_USES_PIPELINED_TEXEL1))
		{
			if (rf & RF_TEX_LOD)
			{
				for (int i = 0; i < 8; i++)
					summary.texel_kinds |= state_tile_kind_bits[i];
			}
			else
			{
				const uint32_t op = (words[0] >> 24) & 63u;
				const uint32_t t = (op == OP_TEX_RECT || op == OP_TEX_RECT_FLIP) ? ((words[1] >> 24) & 7u) : ((words[0] >> 16) & 7u);
				if (rf & (RF_USES_TEXEL0 | RF_USES_PIPELINED_TEXEL1))
					summary.texel_kinds |= state_tile_kind_bits[t];
				if (rf & RF_USES_TEXEL1)
					summary.texel_kinds |= state_tile_kind_bits[(t + 1) & 7u];
			}
		}
	}
	fb.color_write_pending = true;
	if (st.blend_flags & BF_Z_UPDATE)
		fb.depth_write_pending = true;

	const bool full = num_prims == kMaxPrimsPerPass || num_words + 44 + 1024 > kMaxPassWords ||
	                  num_spans + kMaxFbHeight + 2 > kMaxSpansPerPass;
	if (full)
		flush_queues();
}

// A run of back-to-back triangle commands (opcodes 0x08..0x0f), the hot loop of a 3-D frame: 10^8 of them per second
// at the target rate (hard part 0), so everything that does not depend on the individual triangle -- the
// state snapshot, the scissor, which texel kinds its tiles can fetch, where borrowed words go -- is resolved once per
// run and the per-triangle work is: two words read (Y bounds, tile), one 16-byte record written, three counters.
// Returns the number of words consumed (stops at the first command that is not a triangle, or is cut short).
size_t Renderer::draw_triangle_run(const uint32_t *words, size_t num_words)
{
	size_t pos = 0;
	while (pos < num_words)
	{
		const uint32_t op0 = (words[pos] >> 24) & 63u;
		if (op0 < OP_TRI_FILL || op0 > OP_TRI_SHADE_TEX_Z)
			break;
		// Anything unusual goes through the general path, one primitive at a time.
		if (validation || !pass_open || state_dirty)
		{
			const uint32_t len = command_length_words(op0);
			if (pos + len > num_words)
				break;
			draw_primitive(words + pos, len);
			pos += len;
			continue;
		}

		const PassState &st = static_cast<const PassState *>(pass.states)[current_state];
		const uint32_t rf = st.raster_flags;
		const int sc_ylo = int(st.sc_ylo), sc_yhi = int(st.sc_yhi);
		const bool reads_tmem = !(rf & RF_FILL);
		const bool samples = reads_tmem && (rf & (RF_USES_TEXEL0 | RF_USES_TEXEL1 | RF_USES_PIPELINED_TEXEL1)) != 0;
		uint32_t all_tiles = 0;
		for (int i = 0; i < 8; i++)
			all_tiles |= state_tile_kind_bits[i];
		const bool z_update = (st.blend_flags & BF_Z_UPDATE) != 0;
		PrimRecord *records = static_cast<PrimRecord *>(pass.records);
		uint32_t kinds = 0, height = fb.deduced_height;
		uint32_t prims = num_prims, spans = num_spans, jobs = num_jobs;
		bool any = false;
		// Borrowed words (the zero-copy feed): the whole run extends one segment of the caller's buffer, so a triangle's
		// offset in the device word buffer is plain pointer arithmetic.  seg_src == nullptr: copy through place_words.
		const uint32_t *seg_src = nullptr;
		uint32_t seg_dst = 0;
		if (words + pos >= borrowed_base && words + num_words <= borrowed_end)
		{
			const uint32_t first_len = command_length_words(op0);
			if (pos + first_len > num_words)
				break;
			place_words(words + pos, first_len); // opens or extends the segment (the loop below places the same command again: same offset)
			seg_src = segments.back().src;
			seg_dst = segments.back().dst;
		}

		while (pos < num_words)
		{
			const uint32_t *w = words + pos;
			const uint32_t op = (w[0] >> 24) & 63u;
			if (op < OP_TRI_FILL || op > OP_TRI_SHADE_TEX_Z)
				break;
			const uint32_t len = 8u + ((op & 4u) << 2) + ((op & 2u) << 3) + ((op & 1u) << 2);
			if (pos + len > num_words)
				break;
			// The first line of a command a dozen commands ahead (runs are mostly one triangle type: same stride).
			__builtin_prefetch(w + 24u * len);

			const int yl = sign_extend<14>(w[0]), yh = sign_extend<14>(w[1]);
			int line_lo = imax(yh, sc_ylo) >> 2;
			int line_hi = imin(yl - 1, sc_yhi - 1) >> 2;
			if (line_hi - line_lo > 1023)
				line_hi = line_lo + 1023;

			PrimRecord &rec = records[prims];
			if (seg_src)
			{
				rec.word_offset = seg_dst + uint32_t(w - seg_src);
				num_words = rec.word_offset + len;
			}
			else
				rec.word_offset = place_words(w, len);
			rec.span_base = spans;
			rec.state_id = uint16_t(current_state);
			rec.tmem_instance = uint16_t(num_uploads);
			rec.line_lo = int16_t(line_lo);
			rec.line_hi = int16_t(line_hi);
			prims++;
			if (line_hi >= line_lo)
			{
				spans += uint32_t(line_hi - line_lo + 2);
				jobs += span_job_count(line_lo, line_hi);
				height = std::max(height, uint32_t(line_hi + 1));
			}
			if (samples)
			{
				const uint32_t t = (w[0] >> 16) & 7u;
				if (rf & RF_TEX_LOD)
					kinds |= all_tiles;
				else
				{
					if (rf & (RF_USES_TEXEL0 | RF_USES_PIPELINED_TEXEL1))
						kinds |= state_tile_kind_bits[t];
					if (rf & RF_USES_TEXEL1)
						kinds |= state_tile_kind_bits[(t + 1) & 7u];
				}
			}
			pos += len;
			any = true;
			if (prims == kMaxPrimsPerPass || num_words_in_pass() + 44 + 1024 > kMaxPassWords || spans + kMaxFbHeight + 2 > kMaxSpansPerPass)
				break;
		}

		if (seg_src)
			segments.back().count = num_words - seg_dst;
		num_prims = prims;
		num_spans = spans;
		num_jobs = jobs;
		fb.deduced_height = height;
		summary.texel_kinds |= kinds;
		if (any)
		{
			if (reads_tmem)
				pass_reads_tmem = true;
			fb.color_write_pending = true;
			if (z_update)
				fb.depth_write_pending = true;
		}
		if (num_prims == kMaxPrimsPerPass || num_words + 44 + 1024 > kMaxPassWords || num_spans + kMaxFbHeight + 2 > kMaxSpansPerPass)
			flush_queues();
	}
	return pos;
}

void Renderer::note_gpu_write(uint32_t addr, uint32_t bytes)
{
	if (mirror)
		mirror->note_gpu_write(addr, bytes);
}

void Renderer::flush_queues()
{
	if (!pass_open)
		return;

	if (num_prims == 0 && num_uploads == 0)
		return;

	const auto flush_t0 = std::chrono::steady_clock::now();
	struct SubmitBatch
	{
		rdpb_device *d;
		explicit SubmitBatch(rdpb_device *d_) : d(d_) { rdpb_submit_begin(d); }
		~SubmitBatch() { rdpb_submit_end(d); }
	} submit_batch{device};
	rdpb_pass_counts counts = {};
	counts.num_words = num_words;
	counts.num_prims = num_prims;
	counts.num_states = num_states;
	counts.num_uploads = num_uploads;
	counts.num_spans = num_spans;
	rdpb_pass_upload(device, &counts);
	for (const WordSegment &seg : segments)
		rdpb_pass_upload_words(device, seg.src, seg.count, seg.dst);

	const bool need_render = fb.width != 0 && fb.deduced_height != 0 && num_spans != 0;

	// Snapshot 0 must equal live TMEM at pass start; it is stale after a pass with uploads (or a host TMEM poke).
	// A pass that never samples TMEM (fill-only clears) leaves it stale for the next one.
	const bool update_tmem = num_uploads != 0 || (need_render && pass_reads_tmem && !tmem_snapshot_valid);
	if (update_tmem)
		tmem_snapshot_valid = num_uploads == 0;
	// (Decided here because the front launch has to know whether a native span launch follows it.)
	FbInfo probe = {};
	probe.width = fb.width;
	probe.fmt = uint32_t(fb.fmt);
	const bool ssaa_pass = need_render && super_sampled_readback && upscale_log2 != 0 && pass_can_super_sample(probe, upscale_log2);
	const bool native_span_follows = need_render && !ssaa_pass && num_jobs != 0;
	rdpb_launch_front(device, (update_tmem ? 1 : 0) | (uploads_may_read_previous_pass ? 2 : 0) | ((need_render && pass_reads_tmem) ? 4 : 0) |
	                              (native_span_follows ? 8 : 0),
	                  num_uploads, need_render ? num_prims : 0);

	if (need_render)
	{
		FbInfo info = {};
		info.color_addr = fb.addr;
		info.depth_addr = fb.depth_addr;
		info.width = fb.width;
		info.height = std::min<uint32_t>(fb.deduced_height, kMaxFbHeight);
		info.fmt = uint32_t(fb.fmt);
		info.rdram_size = uint32_t(rdram_size);
		info.base_prim_index = base_primitive_index;
		info.num_prims = num_prims;
		info.mask_stride = (num_prims + 31) / 32;
		info.coarse_stride = (info.mask_stride + 31) / 32;
		info.tiles_x = (info.width + kTileW - 1) / kTileW;
		info.tiles_y = (info.height + kTileH - 1) / kTileH;

		info.num_jobs = num_jobs;
		// Internal upscaling (ref rdp_renderer.cpp:2223-2342, 2955-2980): the pass is rendered at native resolution as
		// always -- RDRAM stays bit-exact -- and then once more into the upscaled domain.  Around the native tile launch
		// the upscaled domain picks up whatever changed in native RDRAM behind its back (PRE) and the shadow copy
		// takes the pass's own output (POST).
		// With super-sampled read-back (ref rdp_renderer.cpp:2296-2334, 2955-2980) there is no native pass at all: PRE,
		// the upscaled pass recording which native pixels it stored to, then the resolve that averages those pixels'
		// samples into native RDRAM (and the shadow).
		// Implicit RDRAM coherency: what the CPU may have put into the images this pass reads goes up first (ordered on
		// the compute stream, in front of the tile kernel).
		if (coherency)
		{
			const uint32_t px = info.width * info.height;
			const uint32_t cbpp = fb.fmt == FBFormat::RGBA8888 ? 4 : ((fb.fmt == FBFormat::RGBA5551 || fb.fmt == FBFormat::IA88) ? 2 : 1);
			coherency->gpu_will_read(fb.addr, px * cbpp);
			coherency->gpu_will_read(fb.depth_addr, px * 2);
		}
		const bool ssaa = super_sampled_readback && upscale_log2 != 0 && pass_can_super_sample(info, upscale_log2);
		const bool upscale = ssaa || (upscale_log2 != 0 && pass_wants_upscaling(info, upscale_log2));
		if (!ssaa)
			rdpb_launch_span_bin(device, &info);
		if (upscale)
			rdpb_launch_upscale_sync(device, &info, RDPB_UPSCALE_SYNC_PRE);
		if (!ssaa)
			rdpb_launch_tiles(device, &info, &summary);
		if (upscale)
		{
			if (!ssaa)
				rdpb_launch_upscale_sync(device, &info, RDPB_UPSCALE_SYNC_POST);
			FbInfo up = info;
			up.scale_log2 = upscale_log2;
			up.write_mask = ssaa ? 1u : 0u;
			up.width = info.width << upscale_log2;
			up.height = info.height << upscale_log2;
			up.tiles_x = (up.width + kTileW - 1) / kTileW;
			up.tiles_y = (up.height + kTileH - 1) / kTileH;
			rdpb_launch_span_bin(device, &up);
			rdpb_launch_tiles(device, &up, &summary);
			if (ssaa)
				rdpb_launch_upscale_sync(device, &info, super_sampled_dither ? RDPB_UPSCALE_SYNC_RESOLVE_DITHER : RDPB_UPSCALE_SYNC_RESOLVE);
			upscaled_pass_counter++;
		}

		const uint32_t pixels = info.width * info.height;
		const uint32_t bpp = fb.fmt == FBFormat::RGBA8888 ? 4 : ((fb.fmt == FBFormat::RGBA5551 || fb.fmt == FBFormat::IA88) ? 2 : 1);
		// What this pass renders to: the next pass's texture loads are checked against it (load_tile).
		rendered_ranges[0][0] = fb.addr;
		rendered_ranges[0][1] = fb.addr + pixels * bpp;
		rendered_ranges[1][0] = fb.depth_addr;
		rendered_ranges[1][1] = fb.depth_addr + pixels * 2;
		rendered_ranges_valid = true;
		note_gpu_write(fb.addr, pixels * bpp);
		if (fb.depth_write_pending)
			note_gpu_write(fb.depth_addr, pixels * 2);
	}

	if (!need_render)
		rendered_ranges_valid = false; // nothing rendered: nothing for the next pass's loads to collide with
	uploads_may_read_previous_pass = false;
	rdpb_pass_end(device);
	if (need_render && coherency)
		coherency->pass_submitted();
	pass_counter++;
	flush_ns += uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - flush_t0).count());

	base_primitive_index += num_prims;
	pass_open = false;
	num_words = num_prims = num_states = num_uploads = num_spans = num_jobs = 0;
	staged_words = 0;
	segments.clear();
	pass_reads_tmem = false;
	fb.deduced_height = 0;
	fb.color_write_pending = false;
	fb.depth_write_pending = false;
	state_dirty = true;
}

// ---------------------------------------------------------------------------------------------
// Texture loads
// ---------------------------------------------------------------------------------------------
bool Renderer::tmem_upload_needs_flush(uint32_t addr) const
{
	// Loading from memory a pending pass still has to write: that pass must run first.
	const uint32_t mask = uint32_t(rdram_size - 1);
	const uint32_t pending_pixels = fb.deduced_height * fb.width;
	if (fb.color_write_pending)
	{
		uint32_t offset = (addr - fb.addr) & mask;
		if (fb.fmt == FBFormat::RGBA5551 || fb.fmt == FBFormat::I8)
			offset >>= 1;
		else if (fb.fmt == FBFormat::RGBA8888)
			offset >>= 2;
		if (offset < pending_pixels)
			return true;
	}
	if (fb.depth_write_pending)
	{
		const uint32_t offset = ((addr - fb.depth_addr) & mask) >> 1;
		if (offset < pending_pixels)
			return true;
	}
	return false;
}

void Renderer::load_tile(uint32_t tile, const LoadTileInfo &info)
{
	if (validation && info.mode == UploadMode::TLUT && (info.thi >> 2) > (info.tlo >> 2))
		validation->report_rdp_crash(ValidationError::InvalidMultilineLoadTlut, "Attempting to load multiple lines in TLUT.");

	if (tmem_upload_needs_flush(info.tex_addr))
		flush_queues();
	// The set-up of this pass may run before the previous pass's tile kernel has finished (CUDA backend): say so if
	// this load could read what that pass renders.  Generous bound on the bytes a load can touch: every row up to thi
	// at the image's pitch, or 16 KiB for block / TLUT loads, plus slack.
	if (rendered_ranges_valid && !uploads_may_read_previous_pass)
	{
		const uint64_t row_bytes = (uint64_t(info.tex_width) << unsigned(info.size)) >> 1;
		const uint64_t lo = info.tex_addr & uint32_t(rdram_size - 1);
		const uint64_t hi = lo + (uint64_t(info.thi >> 2) + 2) * row_bytes + 16384 + 4096;
		if (hi > rdram_size)
			uploads_may_read_previous_pass = true;
		for (int r = 0; r < 2; r++)
			if (lo < rendered_ranges[r][1] && rendered_ranges[r][0] < hi)
				uploads_may_read_previous_pass = true;
	}

	// No-op loads.
	if (info.mode != UploadMode::Block)
	{
		if ((info.thi >> 2) < (info.tlo >> 2))
			return;
		if (((((info.shi >> 2) - (info.slo >> 2)) + 1) & 0xfff) == 0)
			return;
	}
	else
	{
		const unsigned pixel_count = ((info.shi - info.slo) + 1) & 0xfff;
		if (!pixel_count || pixel_count > 2048)
			return;
	}

	if (info.mode == UploadMode::Tile)
	{
		const TileDesc &t = regs.tiles[tile];
		const unsigned pixels_per_line = (((info.shi >> 2) - (info.slo >> 2)) + 1) & 0xfff;
		unsigned qwords_per_line = ((pixels_per_line << unsigned(t.size)) + 15) >> 4;
		if (unsigned(t.size) > unsigned(info.size))
			qwords_per_line <<= unsigned(t.size) - unsigned(info.size);
		else if (unsigned(t.size) < unsigned(info.size))
			qwords_per_line >>= unsigned(info.size) - unsigned(t.size);

		const unsigned bytes_per_line = std::max<unsigned>(qwords_per_line * 8, t.stride);
		unsigned max_bytes = 0x1000;
		if (t.fmt == uint8_t(TextureFormat::YUV))
			max_bytes /= 2; // both TMEM halves are written at once
		const unsigned num_lines = ((info.thi >> 2) - (info.tlo >> 2)) + 1;

		if (bytes_per_line * num_lines > max_bytes)
		{
			// More than TMEM holds in one go: wrap-around effects only resolve correctly in <= 4 KiB chunks.
			unsigned lines_per_chunk = (max_bytes / bytes_per_line) & ~1u;
			if (lines_per_chunk == 0)
			{
				fprintf(stderr, "[rdp-b200] LoadTile of more than 2 KiB per line is not supported.\n");
				return;
			}
			for (unsigned line = 0; line < num_lines; line += lines_per_chunk)
			{
				const unsigned n = std::min(num_lines - line, lines_per_chunk);
				LoadTileInfo chunk = info;
				chunk.tlo = info.tlo + (line << 2);
				chunk.thi = chunk.tlo + ((n - 1) << 2);
				load_tile_iteration(tile, chunk, line * t.stride);
			}
			TileDesc &td = regs.tiles[tile];
			td.slo = info.slo;
			td.shi = info.shi;
			td.tlo = info.tlo;
			td.thi = info.thi;
			state_dirty = true;
			return;
		}
	}
	load_tile_iteration(tile, info, 0);
}

void Renderer::load_tile_iteration(uint32_t tile, const LoadTileInfo &info, uint32_t tmem_offset)
{
	TileDesc &t = regs.tiles[tile];
	t.slo = info.slo;
	t.shi = info.shi;
	t.tlo = info.tlo;
	t.thi = info.thi;
	state_dirty = true;

	const unsigned tsize = t.size, vsize = unsigned(info.size);
	const bool yuv = t.fmt == uint8_t(TextureFormat::YUV);

	if (yuv && (tsize != 2 || vsize != 2))
	{
		fprintf(stderr, "[rdp-b200] Only 16bpp is supported for YUV uploads.\n");
		return;
	}
	if (info.size == TextureSize::Bpp4)
	{
		fprintf(stderr, "[rdp-b200] 4-bit VRAM pointer crashes the RDP.\n");
		if (validation)
			validation->report_rdp_crash(ValidationError::LoadTile4bpp, "4-bit VRAM pointer crashes the RDP.");
		return;
	}
	if (tsize == 3 && t.fmt != uint8_t(TextureFormat::RGBA))
	{
		fprintf(stderr, "[rdp-b200] 32bpp tile uploads must use the RGBA texture format.\n");
		return;
	}
	if (info.mode == UploadMode::TLUT && tsize == 3)
	{
		fprintf(stderr, "[rdp-b200] TLUT uploads with 32bpp tiles are unsupported.\n");
		return;
	}
	if (info.mode != UploadMode::TLUT)
	{
		if ((vsize == 3 && tsize == 1) || (vsize == 2 && tsize == 0) || (vsize == 3 && tsize == 0))
		{
			fprintf(stderr, "[rdp-b200] Unsupported texture/tile size mix in load (texture %u, tile %u).\n", vsize, tsize);
			return;
		}
	}

	TmemUpload up = {};
	up.tmem_stride_words = int32_t(t.stride >> 1);
	uint32_t x0 = 0, y0 = 0;
	UploadMode mode = info.mode;

	if (mode == UploadMode::Block)
	{
		x0 = info.slo;
		y0 = info.tlo;

		// LoadBlock gives a texel count and dT/dx; T = floor(x64 * dt / 2048) advances the line.  Decompose it
		// into width/height/stride; when dt does not divide evenly record the candidate-T bounds for the kernel.
		const unsigned pixel_count = (info.shi - info.slo + 1) & 0xfff;
		const unsigned dt = info.thi;
		const unsigned max_iter = (pixel_count - 1) >> (4u - vsize);
		const unsigned max_t = (max_iter * dt) >> 11;

		if (max_t != 0)
		{
			unsigned max_per_line = ((1u << 11) + dt - 1u) / dt;
			unsigned min_per_line = (1u << 11) / dt;
			bool uneven = max_per_line != min_per_line;
			if (uneven)
			{
				unsigned overflow = (dt * max_per_line - (1u << 11)) * max_t;
				if (overflow < dt)
				{
					min_per_line = max_per_line;
					uneven = false;
				}
			}

			up.dxt = int32_t(dt << 5);
			if (tsize == 3 || yuv)
			{
				up.tmem_stride_words <<= 1;
				if (uneven && tsize != vsize)
				{
					fprintf(stderr, "[rdp-b200] LoadBlock with uneven dt and mismatched sizes is unsupported.\n");
					return;
				}
			}

			if (tsize > vsize)
			{
				const unsigned sh = tsize - vsize;
				max_per_line <<= sh;
				min_per_line <<= sh;
				up.dxt >>= sh;
			}
			else if (vsize > tsize)
			{
				fprintf(stderr, "[rdp-b200] LoadBlock: VRAM bpp larger than tile bpp is unsupported.\n");
				return;
			}

			const unsigned max_stride64 = max_per_line + (unsigned(up.tmem_stride_words) >> 2);
			const unsigned min_stride64 = min_per_line + (unsigned(up.tmem_stride_words) >> 2);
			up.min_t_mod = 1.0f / float(max_stride64);
			up.max_t_mod = 1.0f / float(min_stride64);
			up.width = int32_t(pixel_count);
			up.height = 1;
			up.tmem_stride_words >>= 2;
		}
		else
		{
			up.width = int32_t(pixel_count);
			up.height = 1;
			up.tmem_stride_words = 0;
			mode = UploadMode::Tile;
		}
	}
	else
	{
		x0 = info.slo >> 2;
		y0 = info.tlo >> 2;
		up.width = int32_t((((info.shi >> 2) - (info.slo >> 2)) + 1) & 0xfff);
		up.height = int32_t(((info.thi >> 2) - (info.tlo >> 2)) + 1);
	}

	if (!up.width)
		return;

	switch (info.size)
	{
	case TextureSize::Bpp8: up.vram_effective_width = (up.width + 7) & ~7; break;
	case TextureSize::Bpp16: up.vram_effective_width = mode == UploadMode::TLUT ? up.width : ((up.width + 3) & ~3); break;
	case TextureSize::Bpp32: up.vram_effective_width = (up.width + 1) & ~1; break;
	default: break;
	}

	// TMEM is written in 64-bit chunks of four 16-bit words.
	switch (tsize)
	{
	case 0: up.width = ((up.width + 15) & ~15) >> 2; break;
	case 1: up.width = ((up.width + 7) & ~7) >> 1; break;
	case 2:
		up.width = (up.width + 3) & ~3;
		if (yuv)
			up.width >>= 1;
		break;
	default: up.width = (up.width + 1) & ~1; break;
	}

	if (up.height > 1 && mode == UploadMode::TLUT)
	{
		fprintf(stderr, "[rdp-b200] Load TLUT with height > 1 is not supported.\n");
		return;
	}

	up.vram_addr = int32_t(info.tex_addr + ((info.tex_width * y0 + x0) << (vsize - 1)));
	up.vram_width = mode == UploadMode::Block ? up.vram_effective_width : int32_t(info.tex_width);
	up.vram_size = int32_t(vsize);
	up.tmem_offset = int32_t((t.offset + tmem_offset) & 0xfff);
	up.tmem_size = int32_t(tsize);
	up.tmem_fmt = int32_t(t.fmt);
	up.mode = int32_t(mode);
	up.inv_tmem_stride_words = 1.0f / float(up.tmem_stride_words);

	ensure_pass();
	static_cast<TmemUpload *>(pass.uploads)[num_uploads++] = up;
	if (num_uploads >= kMaxUploadsPerPass)
		flush_queues();
}
}
