// rdp_blend.cuh -- the ordered per-pixel back end: RDRAM framebuffer load/store (colour, depth, hidden bits),
// depth compare, blender, coverage accumulation, fill and copy pipes.
//
// Behavioural reference (bit-exact target): parallel-rdp/shaders/memory_interfacing.h:56-571,
// depth_test.h:41-144, z_encode.h:28-56, blender.h:55-143, coverage.h:51-79, dither.h:30-38.
//
// RDRAM layout: host-endian 32-bit words of big-endian N64 memory, so byte address a lives at a^3 and
// 16-bit index i at i^1.  "Hidden" RDRAM holds the 9th bit pair of every 16-bit word, one byte per 16-bit index.
#pragma once
#include "rdp_shade.cuh"

namespace rdpb
{
struct PixelState
{
	int r, g, b, a;   // a carries 3-bit coverage in bits 7..5 for 16-bit formats
	int depth;        // 14-bit compressed Z
	int dz;           // 4-bit compressed delta-Z
	bool color_dirty, depth_dirty;
	uint32_t color_index; // pixel index used by the fill pipe's sub-word select
};

struct FbView
{
	uint8_t *rdram;
	uint8_t *hidden;
	const uint8_t *blend_lut; // blender divider table, 0x8000 entries
	uint32_t fmt;
	uint32_t rdram_size;
	bool alias; // colour and depth image share an address
};

RDP_HD uint16_t *rdram16(uint8_t *p) { return reinterpret_cast<uint16_t *>(p); }
RDP_HD uint32_t *rdram32(uint8_t *p) { return reinterpret_cast<uint32_t *>(p); }

RDP_HD void load_color(const FbView &fb, PixelState &px, uint32_t index)
{
	switch (fb.fmt)
	{
	case FB_I4:
	case FB_I8:
	{
		index &= fb.rdram_size - 1u;
		const int v = fb.rdram[index ^ 3u];
		px.r = px.g = px.b = v;
		px.a = fb.hidden[index >> 1];
		break;
	}
	case FB_RGBA5551:
	{
		index &= (fb.rdram_size >> 1) - 1u;
		const uint32_t w = rdram16(fb.rdram)[index ^ 1u];
		px.r = int((w >> 8) & 0xf8u);
		px.g = int((w >> 3) & 0xf8u);
		px.b = int((w << 2) & 0xf8u);
		px.a = int(((uint32_t(fb.hidden[index]) << 5) | ((w & 1u) << 7)) & 0xffu);
		break;
	}
	case FB_IA88:
	{
		index &= (fb.rdram_size >> 1) - 1u;
		const uint32_t w = rdram16(fb.rdram)[index ^ 1u];
		px.r = px.g = px.b = int(w >> 8);
		px.a = int(w & 0xffu);
		break;
	}
	default: // FB_RGBA8888
	{
		index &= (fb.rdram_size >> 2) - 1u;
		const uint32_t w = rdram32(fb.rdram)[index];
		px.r = int(w >> 24);
		px.g = int((w >> 16) & 0xffu);
		px.b = int((w >> 8) & 0xffu);
		px.a = int(w & 0xffu);
		break;
	}
	}
}

RDP_HD void load_depth(const FbView &fb, PixelState &px, uint32_t index)
{
	index &= (fb.rdram_size >> 1) - 1u;
	const uint32_t w = rdram16(fb.rdram)[index ^ 1u];
	px.depth = int(w >> 2);
	px.dz = int(fb.hidden[index]) | int((w & 3u) << 2);
}

RDP_HD void store_color(const FbView &fb, const PixelState &px, uint32_t index)
{
	if (!px.color_dirty)
		return;
	switch (fb.fmt)
	{
	case FB_I4:
		index &= fb.rdram_size - 1u;
		fb.rdram[index ^ 3u] = 0;
		if (index & 1u)
			fb.hidden[index >> 1] = uint8_t(px.a);
		break;
	case FB_I8:
	{
		index &= fb.rdram_size - 1u;
		const int v = (index & 1u) ? px.g : px.r;
		fb.rdram[index ^ 3u] = uint8_t(v);
		if (index & 1u)
			fb.hidden[index >> 1] = uint8_t((v & 1) * 3);
		break;
	}
	case FB_RGBA5551:
	{
		index &= (fb.rdram_size >> 1) - 1u;
		const uint32_t cov = uint32_t(px.a & 0xff) >> 5;
		const uint32_t w = (uint32_t(px.r & 0xf8) << 8) | (uint32_t(px.g & 0xf8) << 3) | (uint32_t(px.b & 0xf8) >> 2) | (cov >> 2);
		rdram16(fb.rdram)[index ^ 1u] = uint16_t(w);
		fb.hidden[index] = uint8_t(cov & 3u);
		break;
	}
	case FB_IA88:
	{
		index &= (fb.rdram_size >> 1) - 1u;
		const uint32_t w = (uint32_t(px.r & 0xff) << 8) | uint32_t(px.a & 0xff);
		rdram16(fb.rdram)[index ^ 1u] = uint16_t(w);
		fb.hidden[index] = uint8_t((px.a & 1) * 3);
		break;
	}
	default:
	{
		index &= (fb.rdram_size >> 2) - 1u;
		const uint32_t w = (uint32_t(px.r & 0xff) << 24) | (uint32_t(px.g & 0xff) << 16) | (uint32_t(px.b & 0xff) << 8) | uint32_t(px.a & 0xff);
		rdram32(fb.rdram)[index] = w;
		fb.hidden[2u * index] = uint8_t((px.g & 1) * 3);
		fb.hidden[2u * index + 1u] = uint8_t((px.a & 1) * 3);
		break;
	}
	}
}

RDP_HD void store_depth(const FbView &fb, const PixelState &px, uint32_t index)
{
	if (fb.alias || !px.depth_dirty)
		return;
	index &= (fb.rdram_size >> 1) - 1u;
	rdram16(fb.rdram)[index ^ 1u] = uint16_t((uint32_t(px.depth) << 2) | (uint32_t(px.dz) >> 2));
	fb.hidden[index] = uint8_t(px.dz & 3);
}

// When colour and depth share one image every write to one is visible through the other:
// memory_interfacing.h:100-151
RDP_HD void alias_color_to_depth(const FbView &fb, PixelState &px)
{
	if (fb.fmt == FB_RGBA5551)
	{
		px.dz = ((px.a & 0xff) >> 3) | (px.b & 8);
		px.depth = ((px.r & 0xf8) << 6) | ((px.g & 0xf8) << 1) | ((px.b & 0xf8) >> 4);
	}
	else if (fb.fmt == FB_IA88)
	{
		const uint32_t w = (uint32_t(px.r & 0xff) << 8) | uint32_t(px.a & 0xff);
		px.depth = int(w >> 2);
		px.dz = int(((w & 3u) << 2) | ((w & 1u) * 3u));
	}
}

RDP_HD void alias_depth_to_color(const FbView &fb, PixelState &px)
{
	const uint32_t w = (uint32_t(px.depth) << 4) | uint32_t(px.dz);
	if (fb.fmt == FB_RGBA5551)
	{
		px.r = int((w >> 10) & 0xf8u);
		px.g = int((w >> 5) & 0xf8u);
		px.b = int(w & 0xf8u);
		px.a = int((w & 7u) << 5);
	}
	else if (fb.fmt == FB_IA88)
	{
		px.r = int((w >> 10) & 0xffu);
		px.a = int((w >> 2) & 0xffu);
	}
	px.color_dirty = true;
}

RDP_HD void write_color(const FbView &fb, PixelState &px, int r, int g, int b, int a)
{
	px.r = r;
	px.g = g;
	px.b = b;
	if (fb.fmt != FB_I4)
		px.a = a;
	px.color_dirty = true;
}

// Fill pipe: memory_interfacing.h:440-486
RDP_HD void fill_pixel(const FbView &fb, PixelState &px, uint32_t col)
{
	switch (fb.fmt)
	{
	case FB_RGBA8888:
		write_color(fb, px, int(col >> 24), int((col >> 16) & 0xffu), int((col >> 8) & 0xffu), int(col & 0xffu));
		break;
	case FB_RGBA5551:
		col >>= ((px.color_index & 1u) ^ 1u) * 16u;
		write_color(fb, px, int((col >> 8) & 0xf8u), int((col >> 3) & 0xf8u), int((col << 2) & 0xf8u), int((col & 1u) * 0xe0u));
		break;
	case FB_IA88:
	{
		col >>= ((px.color_index & 1u) ^ 1u) * 16u;
		col &= 0xffffu;
		const int i = int(col >> 8);
		write_color(fb, px, i, i, i, int(col & 0xffu));
		break;
	}
	case FB_I8:
	{
		col >>= ((px.color_index & 3u) ^ 3u) * 8u;
		col &= 0xffu;
		write_color(fb, px, int(col), int(col), int(col), int(col));
		break;
	}
	default:
		break;
	}
	if (fb.alias)
		alias_color_to_depth(fb, px);
}

// Copy pipe: memory_interfacing.h:405-438
RDP_HD void copy_pixel(const FbView &fb, PixelState &px, uint32_t word)
{
	switch (fb.fmt)
	{
	case FB_I4:
		px.r = px.g = px.b = px.a = 0;
		px.color_dirty = true;
		break;
	case FB_I8:
		word &= 0xffu;
		write_color(fb, px, int(word), int(word), int(word), int(word));
		break;
	case FB_RGBA5551:
		write_color(fb, px, int((word >> 8) & 0xf8u), int((word >> 3) & 0xf8u), int((word << 2) & 0xf8u), int((word & 1u) * 0xe0u));
		break;
	default:
		break;
	}
	if (fb.alias)
		alias_color_to_depth(fb, px);
}

// ---------------------------------------------------------------------------------------------
// Z encode/decode: z_encode.h:28-56
// ---------------------------------------------------------------------------------------------
RDP_HD int z_decompress(int z)
{
	const int exponent = z >> 11;
	const int mantissa = z & 0x7ff;
	const int shift = imax(6 - exponent, 0);
	const int base = 0x40000 - (0x40000 >> exponent);
	return (mantissa << shift) + base;
}

RDP_HD int z_compress(int z)
{
	const int inv_z = imax(0x3ffff - z, 1);
	const int exponent = iclamp(17 - find_msb(uint32_t(inv_z)), 0, 7);
	const int shift = imax(6 - exponent, 0);
	const int mantissa = (z >> shift) & 0x7ff;
	return (exponent << 11) + mantissa;
}

struct DepthResult
{
	bool pass;
	bool blend_en;
	bool coverage_wrap;
	int shift_a, shift_b;
};

// depth_test.h:41-144
RDP_HD DepthResult depth_test(int z, int dz, int dz_compressed, int mem_depth, int mem_dz_c, int &coverage,
                              int mem_coverage, uint32_t bf, int z_mode)
{
	DepthResult r;
	const bool force_blend = (bf & BF_FORCE_BLEND) != 0;
	const bool aa = (bf & BF_AA) != 0;

	if (bf & BF_Z_COMPARE)
	{
		const int memory_z = z_decompress(mem_depth);
		int memory_dz = 1 << mem_dz_c;
		const int precision = (mem_depth >> 11) & 0xf;
		bool coplanar = false;

		r.shift_a = iclamp(dz_compressed - mem_dz_c, 0, 4);
		r.shift_b = iclamp(mem_dz_c - dz_compressed, 0, 4);

		if (precision < 3)
		{
			if (memory_dz != 0x8000)
				memory_dz = imax(memory_dz << 1, 16 >> precision);
			else
			{
				coplanar = true;
				memory_dz = 0xffff;
			}
		}

		int combined_dz = dz | memory_dz;
		if (combined_dz != 0)
			combined_dz = 1 << find_msb(uint32_t(combined_dz));
		const int combined_dz_ip = combined_dz;
		combined_dz <<= 3;

		const bool farther = coplanar || ((z + combined_dz) >= memory_z);
		const bool overflow = (coverage + mem_coverage) >= 8;
		r.blend_en = force_blend || (!overflow && aa && farther);
		r.coverage_wrap = overflow;

		const bool max_z = memory_z == 0x3ffff;
		const bool front = z < memory_z;
		const bool nearer = coplanar || ((z - combined_dz) <= memory_z);

		switch (z_mode)
		{
		case 0: // opaque
			r.pass = max_z || (overflow ? front : nearer);
			break;
		case 1: // interpenetrating
			if (!front || !farther || !overflow)
				r.pass = max_z || (overflow ? front : nearer);
			else
			{
				const int sh = imax(find_msb(uint32_t(combined_dz_ip & 0xffff)), 0);
				const int cvg_coeff = ((memory_z >> sh) - (z >> sh)) & 0xf;
				coverage = imin((cvg_coeff * coverage) >> 3, 8);
				r.pass = true;
			}
			break;
		case 2: // transparent
			r.pass = front || max_z;
			break;
		default: // decal
			r.pass = farther && nearer && !max_z;
			break;
		}
	}
	else
	{
		r.shift_a = 0;
		r.shift_b = imin(0xf - dz_compressed, 4);
		const bool overflow = (coverage + mem_coverage) >= 8;
		r.blend_en = force_blend || (!overflow && aa);
		r.coverage_wrap = overflow;
		r.pass = true;
	}
	return r;
}

// The depth test exactly as depth_blend_pixel will run it, evaluated before texturing (see EarlyDepth).
RDP_HD bool depth_early_reject(int z, const PrimSetup &p, const EarlyDepth &e, int coverage_count, uint32_t bf, int z_mode)
{
	int coverage = coverage_count;
	const DepthResult d = depth_test(z, p.dz, p.dz_compressed, e.mem_depth, e.mem_dz, coverage, e.mem_coverage, bf, z_mode);
	return !d.pass;
}

// Coverage the blender and the depth test see as "memory coverage": memory_interfacing.h:501-520
RDP_HD int memory_coverage_bits(const FbView &fb, const PixelState &px, uint32_t bf)
{
	if (fb.fmt == FB_I4 || fb.fmt == FB_I8)
		return 0xe0;
	return (bf & BF_IMAGE_READ) ? (px.a & 0xe0) : 0xe0;
}

struct BlendInputs
{
	int pr, pg, pb, pa;     // pixel colour (combiner output, or first blender cycle's output)
	int mr, mg, mb, ma;     // memory colour and coverage (bits 7..5)
	int shade_alpha;
};

// One blender cycle: blender.h:55-143
template <class M>
RDP_HD void blend_cycle(const BlendInputs &in, const PassState &st, uint32_t bf, const M &m, int cycle, const DepthResult &d,
                        bool final_cycle, const uint8_t *lut, int &or_, int &og, int &ob)
{
	const int modes[4] = { m.blend(cycle, 0), m.blend(cycle, 1), m.blend(cycle, 2), m.blend(cycle, 3) };
	const bool force_blend = (bf & BF_FORCE_BLEND) != 0;
	const bool color_on_cvg = (bf & BF_COLOR_ON_CVG) != 0;

	int r1, g1, b1;
	switch (modes[2])
	{
	case 0: r1 = in.pr; g1 = in.pg; b1 = in.pb; break;
	case 1: r1 = in.mr; g1 = in.mg; b1 = in.mb; break;
	case 2: r1 = st.blend_color[0]; g1 = st.blend_color[1]; b1 = st.blend_color[2]; break;
	default: r1 = st.fog_color[0]; g1 = st.fog_color[1]; b1 = st.fog_color[2]; break;
	}

	if (final_cycle && color_on_cvg && !d.coverage_wrap)
	{
		or_ = r1; og = g1; ob = b1;
		return;
	}

	int r0, g0, b0;
	switch (modes[0])
	{
	case 0: r0 = in.pr; g0 = in.pg; b0 = in.pb; break;
	case 1: r0 = in.mr; g0 = in.mg; b0 = in.mb; break;
	case 2: r0 = st.blend_color[0]; g0 = st.blend_color[1]; b0 = st.blend_color[2]; break;
	default: r0 = st.fog_color[0]; g0 = st.fog_color[1]; b0 = st.fog_color[2]; break;
	}

	if (final_cycle && (!d.blend_en || (modes[1] == 0 && modes[3] == 0 && in.pa == 0xff)))
	{
		or_ = r0; og = g0; ob = b0;
		return;
	}

	int a0, a1;
	switch (modes[1])
	{
	case 0: a0 = in.pa; break;
	case 1: a0 = st.fog_color[3]; break;
	case 2: a0 = in.shade_alpha; break;
	default: a0 = 0; break;
	}
	switch (modes[3])
	{
	case 0: a1 = ~a0 & 0xff; break;
	case 1: a1 = in.ma; break;
	case 2: a1 = 0xff; break;
	default: a1 = 0; break;
	}

	a0 >>= 3;
	a1 >>= 3;
	if (modes[3] == 1)
	{
		a0 = (a0 >> d.shift_a) & 0x3c;
		a1 = (a1 >> d.shift_b) | 3;
	}

	int br = r0 * a0 + r1 * (a1 + 1);
	int bg = g0 * a0 + g1 * (a1 + 1);
	int bb = b0 * a0 + b1 * (a1 + 1);

	if (!final_cycle || force_blend)
	{
		or_ = (br >> 5) & 0xff;
		og = (bg >> 5) & 0xff;
		ob = (bb >> 5) & 0xff;
	}
	else
	{
		// Normalising divider, table-driven exactly like the hardware's odd results.
		const int sum = (a0 >> 2) + (a1 >> 2) + 1;
		or_ = lut[(sum << 11) | ((br >> 2) & 0x7ff)];
		og = lut[(sum << 11) | ((bg >> 2) & 0x7ff)];
		ob = lut[(sum << 11) | ((bb >> 2) & 0x7ff)];
	}
}

// dither.h:30-38
RDP_HD int dither_channel(int c, int dith)
{
	const int rounded = c > 247 ? 255 : ((c & 0xf8) + 8);
	const int replace = (dith - (c & 7)) >> 31;
	return (c + ((rounded - c) & replace)) & 0xff;
}

// coverage.h:51-79
RDP_HD int blend_coverage(int coverage, int mem_coverage, bool blend_en, int mode)
{
	switch (mode)
	{
	case 0: return blend_en ? imin(7, mem_coverage + coverage) : ((coverage - 1) & 7);
	case 1: return (coverage + mem_coverage) & 7;
	case 2: return 7;
	default: return mem_coverage;
	}
}

// memory_interfacing.h:488-571, split at the point where colour starts to matter:
//   depth_stage  reads the pixel's depth, delta-Z and memory coverage; decides pass / blend enable / coverage
//   blend_stage  reads the pixel's memory colour and the shaded colour; yields the dithered RGB to store
// depth_blend_pixel runs both back to back (in-order path); the deferred tile kernel runs the depth stage of
// every fragment of a batch first, shades only the survivors, and runs the blend stage afterwards.
struct DepthStage
{
	DepthResult d;
	int coverage;     // coverage count after the depth test (interpenetrating Z modulates it)
	int mem_cvg_bits; // memory coverage in bits 7..5, as the blender's 2B input sees it
};

RDP_HD bool depth_stage(const FbView &fb, const PixelState &px, int z, int dz, int dz_compressed, int coverage, uint32_t bf,
                        int z_mode, DepthStage &o)
{
	o.mem_cvg_bits = memory_coverage_bits(fb, px, bf);
	o.coverage = coverage;
	o.d = depth_test(z, dz, dz_compressed, px.depth, px.dz, o.coverage, o.mem_cvg_bits >> 5, bf, z_mode);
	return o.d.pass && !((bf & BF_AA) && o.coverage == 0);
}

template <class M>
RDP_HD void blend_stage(const FbView &fb, const PixelState &px, const PassState &st, const M &m, const DepthResult &d,
                        int mem_cvg_bits, int sr, int sg, int sb, int sa, int shade_alpha, int dither, int &r, int &g, int &b)
{
	const uint32_t bf = m.bf;
	// Decode what the blender sees as "memory colour".
	BlendInputs in;
	switch (fb.fmt)
	{
	case FB_I4: in.mr = in.mg = in.mb = 0; break;
	case FB_I8: in.mr = in.mg = in.mb = px.r; break;
	case FB_RGBA5551: in.mr = px.r & 0xf8; in.mg = px.g & 0xf8; in.mb = px.b & 0xf8; break;
	case FB_IA88: in.mr = in.mg = in.mb = px.r; break;
	default: in.mr = px.r; in.mg = px.g; in.mb = px.b; break;
	}
	in.ma = mem_cvg_bits;
	in.pr = sr; in.pg = sg; in.pb = sb; in.pa = sa;
	in.shade_alpha = shade_alpha;

	if (bf & BF_TWO_CYCLE)
	{
		blend_cycle(in, st, bf, m, 0, d, false, fb.blend_lut, r, g, b);
		in.pr = r; in.pg = g; in.pb = b;
		blend_cycle(in, st, bf, m, 1, d, true, fb.blend_lut, r, g, b);
	}
	else
		blend_cycle(in, st, bf, m, 0, d, true, fb.blend_lut, r, g, b);

	if (bf & BF_DITHER)
	{
		r = dither_channel(r, dither & 7);
		g = dither_channel(g, (dither >> 3) & 7);
		b = dither_channel(b, (dither >> 6) & 7);
	}
	r &= 0xff;
	g &= 0xff;
	b &= 0xff;
}

template <class M>
RDP_HD void depth_blend_pixel(const FbView &fb, PixelState &px, const PrimSetup &p, const PassState &st,
                              const M &m, const Shaded &sh)
{
	const uint32_t bf = m.bf;
	DepthStage ds;
	if (!depth_stage(fb, px, sh.z, p.dz, p.dz_compressed, sh.coverage, bf, m.z_mode(), ds))
		return;

	int r, g, b;
	blend_stage(fb, px, st, m, ds.d, ds.mem_cvg_bits, sh.r, sh.g, sh.b, sh.a, sh.shade_alpha, sh.dither, r, g, b);

	const int new_coverage = blend_coverage(ds.coverage, ds.mem_cvg_bits >> 5, ds.d.blend_en, m.cvg_mode());
	write_color(fb, px, r, g, b, new_coverage << 5);

	if (bf & BF_Z_UPDATE)
	{
		px.depth = z_compress(sh.z);
		px.dz = p.dz_compressed;
		px.depth_dirty = true;
		if (fb.alias)
			alias_depth_to_color(fb, px);
	}
	else if (fb.alias)
		alias_color_to_depth(fb, px);
}

// ---------------------------------------------------------------------------------------------
// Deferred pipeline: the four per-item steps of the deferred tile kernel (backend_cuda.cu, k_tiles_deferred).
//
// A batch of (pixel, primitive) fragments of one tile is processed as
//   A  candidate : coverage mask + interpolated Z of every fragment          (order-free, compacted over lanes)
//   B  depth     : per pixel, in primitive order: depth test, Z / coverage update; survivors keep a record
//   C  shade     : texture + combiner of the survivors only                   (order-free, compacted over lanes)
//   D  blend     : per pixel, in primitive order: blender against the running memory colour
// which is exact because depth_test / blend_coverage never read colour (depth_test.h:41-144, coverage.h:51-79) and
// shading never reads the framebuffer.  Not valid for states whose coverage depends on the shaded alpha
// (cvg_times_alpha), that can drop a fragment after shading (alpha test), for copy mode, or when the colour and
// depth images alias; the host keeps such passes on the in-order kernel (kRfNotDeferrable).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kRfNotDeferrable = RF_COPY | RF_ALPHA_TEST | RF_CVG_TIMES_ALPHA;

enum DeferBits : uint32_t
{
	// candidate word (phase A -> B)
	DC_Z_MASK = 0x3ffffu,
	DC_COV_SHIFT = 18, // 8-bit sub-sample mask
	DC_VALID = 1u << 26,
	DC_FILL = 1u << 27,
	// survivor record (phase B -> C, D); bits 0..7 hold the sub-sample mask
	DR_BLEND_EN = 1u << 8,
	DR_CVG_WRAP = 1u << 9,
	DR_SHIFT_A_SHIFT = 10, // 3 bits
	DR_SHIFT_B_SHIFT = 13, // 3 bits
	DR_MEM_CVG_SHIFT = 16, // 3 bits
	DR_SURVIVE = 1u << 31,
};

// What phases A and C read of a primitive / span is a subset of the structs; callers fill only those fields.
template <class S>
RDP_HD uint32_t defer_candidate(const PassState &st, const PrimSetup &p, const Span &sp, int x)
{
	if (st.raster_flags & RF_FILL)
		return (x >= sp.start_x && x <= sp.end_x) ? (DC_VALID | DC_FILL) : 0u;
	StaticModes<S> m;
	resolve_modes<S>(st, m);
	const int coverage = shade_coverage(sp, x, m.rf);
	if (coverage == 0)
		return 0u;
	const int z = interpolate_depth(sp, p, x - sp.base_x, coverage);
	return uint32_t(z) | (uint32_t(coverage) << DC_COV_SHIFT) | DC_VALID;
}

// The same for a fragment of a state known not to be fill mode, from the few words phase A of the CUDA tile kernel keeps per
// list entry (whether the state anti-aliases, dz/dx, dz/dy) instead of the state and the primitive themselves.
template <class S>
RDP_HD uint32_t defer_candidate_shaded(bool state_aa, int dzdx, int dzdy, const Span &sp, int x)
{
	const uint32_t rf = ((state_aa ? uint32_t(RF_AA) : 0u) & S::rf_may) | (S::rf_must & uint32_t(RF_AA));
	const int coverage = shade_coverage(sp, x, rf);
	if (coverage == 0)
		return 0u;
	PrimSetup p;
	p.dstzw_dx[2] = dzdx;
	p.dstzw_dy[2] = dzdy;
	const int z = interpolate_depth(sp, p, x - sp.base_x, coverage);
	return uint32_t(z) | (uint32_t(coverage) << DC_COV_SHIFT) | DC_VALID;
}

// Returns the survivor record, 0 when the fragment is rejected.  Updates depth, delta-Z and coverage of the pixel.
// Fill fragments never come here (the caller writes them with fill_pixel and forgets the pixel's earlier survivors).
template <class S>
RDP_HD uint32_t defer_depth(const FbView &fb, PixelState &px, const PassState &st, int dz, int dz_compressed, uint32_t cand)
{
	StaticModes<S> m;
	resolve_modes<S>(st, m);
	const uint32_t bf = m.bf;
	const int z = int(cand & DC_Z_MASK);
	const uint32_t cov8 = (cand >> DC_COV_SHIFT) & 0xffu;
	DepthStage ds;
	if (!depth_stage(fb, px, z, dz, dz_compressed, pop_count(cov8), bf, m.z_mode(), ds))
		return 0u;

	const int new_coverage = blend_coverage(ds.coverage, ds.mem_cvg_bits >> 5, ds.d.blend_en, m.cvg_mode());
	if (fb.fmt != FB_I4)
		px.a = new_coverage << 5;
	px.color_dirty = true;
	if (bf & BF_Z_UPDATE)
	{
		px.depth = z_compress(z);
		px.dz = dz_compressed;
		px.depth_dirty = true;
	}
	return DR_SURVIVE | cov8 | (ds.d.blend_en ? DR_BLEND_EN : 0u) | (ds.d.coverage_wrap ? DR_CVG_WRAP : 0u) |
	       (uint32_t(ds.d.shift_a) << DR_SHIFT_A_SHIFT) | (uint32_t(ds.d.shift_b) << DR_SHIFT_B_SHIFT) |
	       (uint32_t(ds.mem_cvg_bits >> 5) << DR_MEM_CVG_SHIFT);
}

// Shaded colour of a survivor, packed: x = r | g<<8 | b<<16 | a<<24, y = shade_alpha | rgb_dither<<8.
template <class S>
RDP_HD void defer_shade(const ShadeContext &ctx, int x, int y, uint32_t prim, const PrimSetup &p, const Span &sp,
                        const PassState &st, uint32_t record, uint32_t &out_x, uint32_t &out_y)
{
	StaticModes<S> m;
	resolve_modes<S>(st, m);
	const int coverage = int(record & 0xffu);
	int coverage_count = pop_count(uint32_t(coverage));
	ShadedColor c;
	shade_color<S>(ctx, x, y, prim, p, sp, st, m, coverage, coverage_count, c);
	out_x = uint32_t(c.combined.x & 0xff) | (uint32_t(c.combined.y & 0xff) << 8) | (uint32_t(c.combined.z & 0xff) << 16) |
	        (uint32_t(c.combined.w & 0xff) << 24);
	out_y = uint32_t(c.shade_alpha & 0xff) | (uint32_t(c.rgb_dither & 0x1ff) << 8);
}

template <class S>
RDP_HD void defer_blend(const FbView &fb, PixelState &px, const PassState &st, uint32_t record, uint32_t res_x, uint32_t res_y)
{
	StaticModes<S> m;
	resolve_modes<S>(st, m);
	DepthResult d;
	d.pass = true;
	d.blend_en = (record & DR_BLEND_EN) != 0;
	d.coverage_wrap = (record & DR_CVG_WRAP) != 0;
	d.shift_a = int((record >> DR_SHIFT_A_SHIFT) & 7u);
	d.shift_b = int((record >> DR_SHIFT_B_SHIFT) & 7u);
	int r, g, b;
	blend_stage(fb, px, st, m, d, int((record >> DR_MEM_CVG_SHIFT) & 7u) << 5, int(res_x & 0xffu), int((res_x >> 8) & 0xffu),
	            int((res_x >> 16) & 0xffu), int(res_x >> 24), int(res_y & 0xffu), int(res_y >> 8), r, g, b);
	px.r = r;
	px.g = g;
	px.b = b;
}

// Walk one pixel through one binned primitive (shade, then fill / copy / depth-blend).
// Returns true when the pixel was shaded (a "shaded pixel" of the Mpix/s metric: shade_pixel returned true).
// Fill-mode states are handled here for every specialisation; everything else goes through Spec S.
// The part after the primitive / span / state look-ups (the CUDA tile kernel resolves those warp-cooperatively).
template <class S>
RDP_HD bool process_pixel_span(const ShadeContext &ctx, const FbView &fb, PixelState &px, int x, int y, uint32_t prim,
                               const PrimSetup &p, const Span &sp, const PassState &st)
{
	if (st.raster_flags & RF_FILL)
	{
		if (x < sp.start_x || x > sp.end_x)
			return false;
		fill_pixel(fb, px, st.fill_color);
		return true;
	}

	StaticModes<S> m;
	resolve_modes<S>(st, m);
	Shaded sh;
	EarlyDepth early;
	early.mem_depth = px.depth;
	early.mem_dz = px.dz;
	early.mem_coverage = memory_coverage_bits(fb, px, m.bf) >> 5;
	const int shaded = shade_pixel<S>(ctx, x, y, prim, p, sp, st, m, early, sh);
	if (shaded != 1)
		return shaded == 2; // a depth-rejected fragment still counts as shaded for the Mpix/s metric (the reference shades it)
	if (sh.kind == 2)
		copy_pixel(fb, px, sh.raw);
	else
		depth_blend_pixel(fb, px, p, st, m, sh);
	return true;
}

template <class S>
RDP_HD bool process_pixel_primitive(const ShadeContext &ctx, const FbView &fb, PixelState &px, int x, int y, uint32_t prim)
{
	const PrimSetup &p = ctx.prims[prim];
	const int sl = ctx.scale_log2;
	const int line_lo = int(p.line_lo) << sl, line_hi = (int(p.line_hi) << sl) + ((1 << sl) - 1);
	if (y < line_lo || y > line_hi)
		return false;
	const Span &sp = ctx.spans[(p.span_base << sl) + uint32_t(y - line_lo)];
	if (!sp.valid)
		return false;
	// Not upscaled: every sub-pixel is shaded as its native pixel's first sub-pixel (shading.h:45-52).  The span stored
	// for line y already is the snapped line's (compute_span).
	if (sl != 0 && (p.flags & SF_DISABLE_UPSCALING))
	{
		x &= ~((1 << sl) - 1);
		y &= ~((1 << sl) - 1);
	}
	return process_pixel_span<S>(ctx, fb, px, x, y, prim, p, sp, ctx.states[p.state_id]);
}

// Blender divider table (reference data: parallel-rdp/luts.hpp blender_lut, itself generated by the oracle's
// blender_init_lut, angrylion rdp/blender.c:304).  Regenerated here from the rule rather than stored:
// entry[(d << 11) | n] = the RDP's 8-bit quotient of n/d with its characteristic truncation.
void build_blender_lut(uint8_t *lut); // defined in rdp_luts.cpp
}
