// rdp_shade.cuh -- per-pixel shading: coverage, attribute interpolation, perspective divide, LOD,
// TMEM texel fetch + filtering, colour combiner, alpha test, dither coefficients.
//
// Behavioural reference (bit-exact target): parallel-rdp/shaders/shading.h:38-359 and its callees
//   coverage.h:31-44, interpolation.h:30-142, perspective.h:27-112, clamping.h:40-76, texture.h:34-1003,
//   combiner.h:28-282, dither.h:40-68, noise.h:29-69.
// Everything is 32-bit integer arithmetic; the code is host+device so the same source is checked on the CPU
// against the Angrylion oracle (tests/hostsim) and runs in the CUDA tile kernel.
#pragma once
#include "rdp_setup.cuh"
#include "rdp_spec.h"

namespace rdpb
{
struct Vec4
{
	int x, y, z, w;
};

RDP_HD Vec4 vec4(int x, int y, int z, int w)
{
	Vec4 v;
	v.x = x;
	v.y = y;
	v.z = z;
	v.w = w;
	return v;
}
RDP_HD Vec4 splat4(int v) { return vec4(v, v, v, v); }

// What the shading stage hands to the blender for one (pixel, primitive).
struct Shaded
{
	int r, g, b, a;
	int z;           // 18-bit depth
	int dither;      // 9-bit rgb dither pattern
	int coverage;    // 0..8 sub-samples
	int shade_alpha; // shade alpha + alpha dither, for the blender's 1B input
	uint32_t raw;    // copy pipe: raw texel word
	int kind;        // 0 = normal, 1 = fill, 2 = copy
};

// What the pixel currently holds in the depth buffer, handed to shade_pixel so that a fragment the depth test is
// going to reject skips texturing and the combiner (the result is identical: a rejected fragment writes nothing;
// depth_test.h:41-144 only needs Z, delta-Z and the coverage count, none of which depend on colour unless coverage
// is modulated by alpha, in which case the early test is not attempted).
struct EarlyDepth
{
	int mem_depth, mem_dz, mem_coverage;
};

// Read-only view of everything a pass's pixels consume.
struct ShadeContext
{
	const PrimSetup *prims;
	const PassState *states;
	const Span *spans;
	const uint8_t *tmem; // TMEM snapshots, kTmemBytes each
	uint32_t fb_fmt;
	uint32_t base_prim_index;
	// log2 of the internal upscaling factor of this pass (0 at native resolution; a literal 0 in every kernel that is
	// not the upscaled instantiation, so the extra shifts fold away).  Pixel coordinates, span indices and every X in a
	// Span are then in upscaled pixels; ref shaders/shading.h:27-28, interpolation.h (SCALING_LOG2).
	int scale_log2;
};

// ---------------------------------------------------------------------------------------------
// Noise (deterministic hash, as in the reference fork of the oracle): noise.h:29-69
// ---------------------------------------------------------------------------------------------
RDP_HD uint32_t noise_hash(uint32_t x, uint32_t y, uint32_t prim)
{
	const uint32_t P = 1103515245u;
	uint32_t a = x, b = y, c = prim;
	for (int i = 0; i < 3; i++)
	{
		uint32_t na = ((a >> 8) ^ b) * P;
		uint32_t nb = ((b >> 8) ^ c) * P;
		uint32_t nc = ((c >> 8) ^ a) * P;
		a = na;
		b = nb;
		c = nc;
	}
	return (a >> 16) & 0xffffu;
}

// ---------------------------------------------------------------------------------------------
// Clamps: clamping.h:40-76
// ---------------------------------------------------------------------------------------------
RDP_HD int special_9bit_expand(int v)
{
	// 9-bit sign extension around 0x80 without clamping.
	return sign_extend<9>(uint32_t(v - 0x80)) + 0x80;
}

RDP_HD int clamp_9bit(int v)
{
	return iclamp(special_9bit_expand(v), 0, 0xff);
}

RDP_HD int clamp_depth(int z)
{
	z -= 1 << 17;
	z = sign_extend<19>(uint32_t(z));
	z += 1 << 17;
	return iclamp(z, 0, 0x3ffff);
}

// ---------------------------------------------------------------------------------------------
// Coverage: coverage.h:31-44.  Bit layout: two samples per sub-scanline, offsets {0,4},{2,6},{0,4},{2,6}.
// ---------------------------------------------------------------------------------------------
RDP_HD int compute_coverage_scalar(const Span &s, int x)
{
	const int base = (x << 3) & 0xffff;
	int cov = 0;
	for (int line = 0; line < 4; line++)
	{
		const int l = s.xl[line], r = s.xr[line];
		const int o0 = (base + ((line & 1) ? 2 : 0)) & 0xffff;
		const int o1 = (base + ((line & 1) ? 6 : 4)) & 0xffff;
		if (o0 >= l && o0 < r)
			cov |= 1 << (2 * line);
		if (o1 >= l && o1 < r)
			cov |= 2 << (2 * line);
	}
	return cov;
}

// The same eight comparisons two sub-scanlines at a time, in the 16-bit halves of a word.  Sample positions stay
// below 0x7fff for x < 4095, and the edges are either valid x.3 values below 0x8000 or the invalid-line sentinels
// (left 0xffff, right 0), which masking bit 15 off turns into "never reached" / "never below" -- so
// (P | 0x8000) - E cannot borrow across the halves and its bit 15 says P >= E.  With 8x internal upscaling edges
// reach beyond 0x7fff (scissor 4095 * 2 * 8): `wide` keeps those passes on the scalar form.
RDP_HD int compute_coverage(const Span &s, int x, bool wide = false)
{
	if (wide || uint32_t(x) >= 4095u)
		return compute_coverage_scalar(s, x); // (also the wrapped positions of a negative x)
	const uint32_t H = 0x80008000u;
	const uint32_t base = uint32_t(x) << 3;
	const uint32_t p0 = (base | ((base + 2u) << 16)) | H; // samples {0, 2} of an even / odd sub-scanline pair
	const uint32_t p1 = ((base + 4u) | ((base + 6u) << 16)) | H; // samples {4, 6}
	uint32_t cov = 0;
	for (int pair = 0; pair < 2; pair++)
	{
		const uint32_t l = (uint32_t(s.xl[2 * pair]) | (uint32_t(s.xl[2 * pair + 1]) << 16)) & ~H;
		const uint32_t r = (uint32_t(s.xr[2 * pair]) | (uint32_t(s.xr[2 * pair + 1]) << 16)) & ~H;
		const uint32_t c0 = (p0 - l) & ~(p0 - r) & H; // bit 15: first line of the pair, bit 31: second
		const uint32_t c1 = (p1 - l) & ~(p1 - r) & H;
		const uint32_t m = (c0 >> 15) | (c1 >> 14);   // bits 0, 1: first line; bits 16, 17: second
		cov |= ((m | (m >> 14)) & 0xfu) << (4 * pair);
	}
	return int(cov);
}

// ---------------------------------------------------------------------------------------------
// Perspective divide: perspective.h:27-112
// ---------------------------------------------------------------------------------------------
// 64 x (reciprocal base, slope) pairs of the RDP's normalised-W reciprocal ROM, packed base | slope << 16.
// Indexed per lane, so on the device it lives in global memory (L1-resident) rather than in the constant bank
// (divergent constant reads serialise) or on the stack (rebuilt per call).
#define RDP_PERSP_ROM_INIT                                                                                              \
	{                                                                                                                   \
		RDP_PR(0x4000, -252), RDP_PR(0x3f04, -244), RDP_PR(0x3e10, -238), RDP_PR(0x3d22, -230), RDP_PR(0x3c3c, -223),   \
		RDP_PR(0x3b5d, -218), RDP_PR(0x3a83, -210), RDP_PR(0x39b1, -205), RDP_PR(0x38e4, -200), RDP_PR(0x381c, -194),   \
		RDP_PR(0x375a, -189), RDP_PR(0x369d, -184), RDP_PR(0x35e5, -179), RDP_PR(0x3532, -175), RDP_PR(0x3483, -170),   \
		RDP_PR(0x33d9, -166), RDP_PR(0x3333, -162), RDP_PR(0x3291, -157), RDP_PR(0x31f4, -155), RDP_PR(0x3159, -150),   \
		RDP_PR(0x30c3, -147), RDP_PR(0x3030, -143), RDP_PR(0x2fa1, -140), RDP_PR(0x2f15, -137), RDP_PR(0x2e8c, -134),   \
		RDP_PR(0x2e06, -131), RDP_PR(0x2d83, -128), RDP_PR(0x2d03, -125), RDP_PR(0x2c86, -123), RDP_PR(0x2c0b, -120),   \
		RDP_PR(0x2b93, -117), RDP_PR(0x2b1e, -115), RDP_PR(0x2aab, -113), RDP_PR(0x2a3a, -110), RDP_PR(0x29cc, -108),   \
		RDP_PR(0x2960, -106), RDP_PR(0x28f6, -104), RDP_PR(0x288e, -102), RDP_PR(0x2828, -100), RDP_PR(0x27c4, -98),    \
		RDP_PR(0x2762, -96), RDP_PR(0x2702, -94), RDP_PR(0x26a4, -92), RDP_PR(0x2648, -91), RDP_PR(0x25ed, -89),        \
		RDP_PR(0x2594, -87), RDP_PR(0x253d, -86), RDP_PR(0x24e7, -85), RDP_PR(0x2492, -83), RDP_PR(0x243f, -81),        \
		RDP_PR(0x23ee, -80), RDP_PR(0x239e, -79), RDP_PR(0x234f, -77), RDP_PR(0x2302, -76), RDP_PR(0x22b6, -74),        \
		RDP_PR(0x226c, -74), RDP_PR(0x2222, -72), RDP_PR(0x21da, -71), RDP_PR(0x2193, -70), RDP_PR(0x214d, -69),        \
		RDP_PR(0x2108, -67), RDP_PR(0x20c5, -67), RDP_PR(0x2082, -65), RDP_PR(0x2041, -65),                             \
	}
#define RDP_PR(base, slope) (uint32_t(base) | (uint32_t(uint16_t(int16_t((slope) * 4))) << 16))
#if defined(__CUDACC__)
static __device__ const uint32_t g_persp_rom_device[64] = RDP_PERSP_ROM_INIT;
#endif
static const uint32_t g_persp_rom_host[64] = RDP_PERSP_ROM_INIT;

RDP_HD uint32_t persp_rom(int i)
{
#if defined(__CUDA_ARCH__)
	return g_persp_rom_device[i];
#else
	return g_persp_rom_host[i];
#endif
}

RDP_HD void perspective_divide(int s, int t, int w, int &os, int &ot, bool &overflow)
{
	const bool w_carry = w <= 0;
	w &= 0x7fff;

	const int shift = imin(14 - find_msb(uint32_t(w)), 14);
	const int normout = (w << shift) & 0x3fff;
	const int wnorm = normout & 0xff;
	const int idx = normout >> 8;
	const uint32_t rom = persp_rom(idx);
	const int rcp = ((int(int16_t(rom >> 16)) * wnorm) >> 10) + int(rom & 0xffffu);

	int prod[2] = { s * rcp, t * rcp };
	const int temp_mask = ((1 << 30) - 1) & -((1 << 29) >> shift);
	int res[2];
	for (int i = 0; i < 2; i++)
	{
		const int oob = prod[i] & temp_mask;
		int v;
		if (shift != 14)
		{
			prod[i] = prod[i] >> (13 - shift);
			v = prod[i];
		}
		else
			v = int(uint32_t(prod[i]) << 1);

		if (oob != temp_mask && oob != 0)
		{
			v = (prod[i] & (1 << 29)) == 0 ? 0x7fff : -0x8000;
			overflow = true;
		}
		res[i] = v;
	}

	if (w_carry)
	{
		res[0] = res[1] = 0x7fff;
		overflow = true;
	}

	os = iclamp(res[0], -0x10000, 0xffff);
	ot = iclamp(res[1], -0x10000, 0xffff);
}

RDP_HD void divide_st(bool perspective, int s, int t, int w, int &os, int &ot, bool &overflow)
{
	if (perspective)
		perspective_divide(s, t, w, os, ot, overflow);
	else
	{
		os = s;
		ot = t;
	}
}

// ---------------------------------------------------------------------------------------------
// Attribute interpolation: interpolation.h:30-142
// ---------------------------------------------------------------------------------------------
RDP_HD void centroid_offset(int coverage, int &xoff, int &yoff)
{
	const int first = find_lsb(uint32_t(coverage));
	yoff = first >> 1;
	xoff = ((first & 1) << 1) + (yoff & 1);
}

RDP_HD Vec4 interpolate_shade(const Span &sp, const PrimSetup &p, int dx, int coverage, int sl = 0)
{
	int xoff, yoff;
	centroid_offset(coverage, xoff, yoff);
	int out[4];
	for (int i = 0; i < 4; i++)
	{
		int c = int(uint32_t(sp.rgba[i]) + uint32_t((p.drgba_dx[i] & ~0x1f) >> sl) * uint32_t(dx));
		int snapped = c >> 14;
		snapped = int(uint32_t(snapped) << (2 + sl));
		snapped += xoff * (p.drgba_dx[i] >> 14) + yoff * (p.drgba_dy[i] >> 14);
		snapped >>= 4 + sl;
		out[i] = clamp_9bit(snapped);
	}
	return vec4(out[0], out[1], out[2], out[3]);
}

RDP_HD int interpolate_depth(const Span &sp, const PrimSetup &p, int dx, int coverage, int sl = 0)
{
	int xoff, yoff;
	centroid_offset(coverage, xoff, yoff);
	// Whole native pixels step by dz/dx, the sub-pixel remainder by dz/dx quantised to the scale (interpolation.h:131).
	int z = int(uint32_t(sp.stzw[2]) + uint32_t(p.dstzw_dx[2]) * uint32_t(dx >> sl) +
	            uint32_t(p.dstzw_dx[2] >> sl) * uint32_t(dx & ((1 << sl) - 1)));
	int snapped = z >> 10;
	snapped = int(uint32_t(snapped) << (2 + sl));
	snapped += xoff * (p.dstzw_dx[2] >> 10) + yoff * (p.dstzw_dy[2] >> 10);
	snapped >>= 5 + sl;
	return clamp_depth(snapped);
}

// ---------------------------------------------------------------------------------------------
// TMEM access.  TMEM is stored like RDRAM: host-endian 16-bit words of big-endian memory,
// so byte address a lives at a^3 and 16-bit index i at i^1 (texture.h:98-100,193-195).
// ---------------------------------------------------------------------------------------------
RDP_HD uint32_t tmem_read8(const uint8_t *tmem, uint32_t idx) { return tmem[idx & 0xfffu]; }
RDP_HD uint32_t tmem_read16(const uint8_t *tmem, uint32_t idx)
{
	return reinterpret_cast<const uint16_t *>(tmem)[idx & 0x7ffu];
}

RDP_HD Vec4 expand_rgba16(uint32_t word)
{
	uint32_t r = (word >> 11) & 31u, g = (word >> 6) & 31u, b = (word >> 1) & 31u;
	r = (r << 3) | (r >> 2);
	g = (g << 3) | (g >> 2);
	b = (b << 3) | (b >> 2);
	return vec4(int(r), int(g), int(b), int((word & 1u) * 0xffu));
}

RDP_HD Vec4 expand_ia16(uint32_t word)
{
	const int i = int(word >> 8), a = int(word & 0xff);
	return vec4(i, i, i, a);
}

// Fetch one texel.  lut_bank / addr_xor select which of the four replicated TLUT banks is read
// (texture.h:150-171, :556-600).  chroma_x only matters for YUV.
RDP_HD Vec4 fetch_texel(int kind, const TileDesc &tile, const uint8_t *tmem, int s, int t, uint32_t lut_bank,
                        uint32_t addr_xor, bool tlut_ia, int chroma_x)
{
	const uint32_t us = uint32_t(s), ut = uint32_t(t);
	const uint32_t row = tile.offset + tile.stride * ut;
	const uint32_t swz8 = ((ut & 1u) << 2) ^ 3u;
	const uint32_t swz16 = ((ut & 1u) << 1) ^ 1u;

	switch (kind)
	{
	case TK_I4:
	case TK_IA4:
	case TK_CI4:
	case TK_CI4_TLUT:
	{
		const uint32_t mask = kind == TK_CI4_TLUT ? 0x7ffu : 0xfffu;
		const uint32_t idx = ((row + (us >> 1)) & mask) ^ swz8;
		uint32_t nib = (tmem_read8(tmem, idx) >> ((~us & 1u) * 4u)) & 0xfu;
		if (kind == TK_I4)
		{
			nib |= nib << 4;
			return splat4(int(nib));
		}
		if (kind == TK_IA4)
		{
			uint32_t i = nib & 0xeu;
			i = (i << 4) | (i << 1) | (i >> 2);
			return vec4(int(i), int(i), int(i), int((nib & 1u) * 0xffu));
		}
		nib |= uint32_t(tile.palette) << 4;
		if (kind == TK_CI4)
			return splat4(int(nib));
		const uint32_t entry = ((nib << 2) + lut_bank) ^ addr_xor;
		const uint32_t word = tmem_read16(tmem, 0x400u | entry);
		return tlut_ia ? expand_ia16(word) : expand_rgba16(word);
	}

	case TK_I8:
	case TK_IA8:
	case TK_CI8_TLUT:
	{
		const uint32_t mask = kind == TK_CI8_TLUT ? 0x7ffu : 0xfffu;
		const uint32_t idx = ((row + us) & mask) ^ swz8;
		const uint32_t byte = tmem_read8(tmem, idx);
		if (kind == TK_I8)
			return splat4(int(byte));
		if (kind == TK_IA8)
		{
			uint32_t i = byte >> 4, a = byte & 0xfu;
			i |= i << 4;
			a |= a << 4;
			return vec4(int(i), int(i), int(i), int(a));
		}
		const uint32_t entry = ((byte << 2) + lut_bank) ^ addr_xor;
		const uint32_t word = tmem_read16(tmem, 0x400u | entry);
		return tlut_ia ? expand_ia16(word) : expand_rgba16(word);
	}

	case TK_RGBA16:
	case TK_IA16:
	case TK_RAW16:
	case TK_CI16_TLUT:
	{
		const uint32_t mask = kind == TK_CI16_TLUT ? 0x7ffu : 0xfffu;
		const uint32_t idx = (((row + us * 2u) & mask) >> 1) ^ swz16;
		const uint32_t word = tmem_read16(tmem, idx);
		if (kind == TK_RGBA16)
			return expand_rgba16(word);
		if (kind == TK_IA16)
			return expand_ia16(word);
		if (kind == TK_RAW16)
			return vec4(int(word >> 8), int(word & 0xff), int(word >> 8), int(word & 0xff));
		const uint32_t entry = (((word >> 6) & ~3u) + lut_bank) ^ addr_xor;
		const uint32_t lut = tmem_read16(tmem, 0x400u | entry);
		return tlut_ia ? expand_ia16(lut) : expand_rgba16(lut);
	}

	case TK_RGBA32:
	{
		const uint32_t idx = (((row + us * 2u) & 0x7ffu) >> 1) ^ swz16;
		const uint32_t lo = tmem_read16(tmem, idx);
		const uint32_t hi = tmem_read16(tmem, idx | 0x400u);
		return vec4(int(lo >> 8), int(lo & 0xff), int(hi >> 8), int(hi & 0xff));
	}

	default: // TK_YUV16
	{
		const uint32_t luma_idx = ((row + us) & 0x7ffu) ^ swz8;
		const uint32_t chroma_idx = (((row + uint32_t(chroma_x) * 2u) & 0x7ffu) >> 1) ^ swz16;
		const int luma = int(tmem_read8(tmem, luma_idx | 0x800u));
		const uint32_t chroma = tmem_read16(tmem, chroma_idx);
		return vec4(int(chroma >> 8) - 0x80, int(chroma & 0xff) - 0x80, luma, luma);
	}
	}
}

// Coordinate wrap/mirror/mask: texture.h:34-73
RDP_HD int mask_coord(int c, int mask_bits, bool mirror)
{
	if (mask_bits != 0)
	{
		const int m = 1 << mask_bits;
		if (mirror)
			c ^= imax((c & m) - 1, 0);
		c &= m - 1;
	}
	return c;
}

// texture.h:317-360
RDP_HD int shift_texcoord(int coord, int shift)
{
	coord = iclamp(coord, -0x8000, 0x7fff);
	if (shift < 11)
		coord >>= shift;
	else
	{
		coord = int(uint32_t(coord) << (32 - shift));
		coord >>= 16;
	}
	return coord;
}

RDP_HD int clamp_shift_texcoord(bool clamp_bit, int coord, int lo, int hi, int shift)
{
	coord = shift_texcoord(coord, shift);
	if (clamp_bit)
	{
		if ((coord >> 3) >= hi)
			coord = (((hi >> 2) - (lo >> 2)) & 0x3ff) << 5;
		else
			coord = imax(coord - (lo << 3), 0);
	}
	else
		coord -= lo << 3;
	return coord;
}

// YUV / "convert" stage applied when a cycle does not bilerp: texture.h:487-496
RDP_HD Vec4 texture_convert(const Vec4 &in, const int16_t *k)
{
	const int r = sign_extend<9>(uint32_t(in.x)), g = sign_extend<9>(uint32_t(in.y)), b = sign_extend<9>(uint32_t(in.z));
	return vec4(b + ((k[0] * g + 0x80) >> 8), b + ((k[1] * r + k[2] * g + 0x80) >> 8), b + ((k[3] * r + 0x80) >> 8), b);
}

RDP_HD int bilerp3(int base, int t10, int t01, int f0, int f1)
{
	return (((t10 - base) * f0 + (t01 - base) * f1 + 0x10) >> 5) + base;
}

// Full texture sample for one cycle: texture.h:498-923
// TK_MASK: bit k set when texel kind k can occur (~0u in the generic kernel).
template <uint32_t TK_MASK>
RDP_HD Vec4 sample_texture(const TileDesc &tile, const uint8_t *tmem, int s, int t, uint32_t rf, bool bilerp,
                           bool convert_one, const int16_t *k, const Vec4 &prev_cycle)
{
	constexpr bool kind_is_static = ct_single_bit(TK_MASK);
	constexpr int static_kind = int(ct_log2(TK_MASK));
	constexpr bool may_be_yuv = (TK_MASK & (1u << TK_YUV16)) != 0;
	const bool tlut = (rf & RF_TLUT) != 0;
	const bool tlut_ia = (rf & RF_TLUT_IA) != 0;
	const bool sample_quad = (rf & RF_SAMPLE_QUAD) != 0;
	const bool mid_texel_state = (rf & RF_MID_TEXEL) != 0;

	s = clamp_shift_texcoord((tile.flags & TF_CLAMP_S) != 0, s, int(tile.slo), int(tile.shi), tile.shift_s);
	t = clamp_shift_texcoord((tile.flags & TF_CLAMP_T) != 0, t, int(tile.tlo), int(tile.thi), tile.shift_t);

	int frac_s = 0, frac_t = 0;
	if (sample_quad || tlut)
	{
		frac_s = s & 31;
		frac_t = t & 31;
	}
	int sum_frac = frac_s + frac_t;
	s >>= 5;
	t >>= 5;

	int s0 = mask_coord(s, tile.mask_s, (tile.flags & TF_MIRROR_S) != 0);
	int t0 = mask_coord(t, tile.mask_t, (tile.flags & TF_MIRROR_T) != 0);
	int s1 = mask_coord(s + 1, tile.mask_s, (tile.flags & TF_MIRROR_S) != 0);
	int t1 = mask_coord(t + 1, tile.mask_t, (tile.flags & TF_MIRROR_T) != 0);

	// T only carries 8 bits between the two rows.
	const int tdiff = imax(t1 - t0, -255);
	t1 = (t0 & 0xff) + tdiff;
	t0 &= 0xff;

	const bool mid_texel = mid_texel_state && bilerp && frac_s == 0x10 && frac_t == 0x10;
	const bool upper_lut = sum_frac >= 0x20;
	if (mid_texel)
		sum_frac = 0; // forces all four texels to be fetched

	const bool yuv = may_be_yuv && !tlut && tile.fmt == 1;
	const bool upper = sum_frac >= 0x20;
	int base_s = upper ? s1 : s0;
	int base_t = upper ? t1 : t0;
	const int chroma_frac = ((s0 & 1) << 4) | (frac_s >> 1);

	const int kind = kind_is_static ? static_kind : resolve_texel_kind(tile.fmt, tile.size, tlut);
	Vec4 tb, t10 = splat4(0), t01 = splat4(0), t11 = splat4(0);

	if (tlut)
	{
		if (!sample_quad)
		{
			// Bilinear footprint over the four TLUT banks of one texel.
			base_s = s0;
			base_t = t0;
			s1 = s0;
			t1 = t0;
		}
		const uint32_t addr_xor = upper_lut ? 2u : 1u;
		tb = fetch_texel(kind, tile, tmem, base_s, base_t, upper ? 3u : 0u, addr_xor, tlut_ia, 0);
		if (bilerp)
		{
			t10 = fetch_texel(kind, tile, tmem, s1, t0, 1u, addr_xor, tlut_ia, 0);
			t01 = fetch_texel(kind, tile, tmem, s0, t1, 2u, addr_xor, tlut_ia, 0);
		}
		if (mid_texel)
			t11 = fetch_texel(kind, tile, tmem, s1, t1, 3u, addr_xor, tlut_ia, 0);
	}
	else if (yuv)
	{
		const int chroma_x0 = int(uint32_t(s0) >> 1);
		const int chroma_x1 = int(uint32_t(s1 + (s1 - s0)) >> 1);
		tb = fetch_texel(kind, tile, tmem, s0, t0, 0, 0, false, chroma_x0);
		if (sample_quad)
		{
			t10 = fetch_texel(kind, tile, tmem, s1, t0, 0, 0, false, chroma_x1);
			t01 = fetch_texel(kind, tile, tmem, s0, t1, 0, 0, false, chroma_x0);
			t11 = fetch_texel(kind, tile, tmem, s1, t1, 0, 0, false, chroma_x1);
		}
	}
	else
	{
		tb = fetch_texel(kind, tile, tmem, base_s, base_t, 0, 0, false, 0);
		if (sample_quad)
		{
			t10 = fetch_texel(kind, tile, tmem, s1, t0, 0, 0, false, 0);
			t01 = fetch_texel(kind, tile, tmem, s0, t1, 0, 0, false, 0);
		}
		if (mid_texel)
			t11 = fetch_texel(kind, tile, tmem, s1, t1, 0, 0, false, 0);
	}

	Vec4 acc;
	if (convert_one)
	{
		// bilerp + convert_one: the previous cycle's texel supplies the interpolation factors.
		const int pr = sign_extend<9>(uint32_t(prev_cycle.x));
		const int pg = sign_extend<9>(uint32_t(prev_cycle.y));
		const int pb = sign_extend<9>(uint32_t(prev_cycle.z));
		if (sample_quad)
		{
			const bool mid_rg = yuv ? (mid_texel_state && chroma_frac == 0x10 && frac_t == 0x10) : mid_texel;
			const bool mid_ba = mid_texel;
			const bool upper_ba = sum_frac >= 32;
			const bool upper_rg = yuv ? ((chroma_frac + frac_t) >= 32 && !mid_rg) : upper_ba;

			const int f_rg0 = upper_rg ? pg : pr, f_rg1 = upper_rg ? pr : pg;
			const int f_ba0 = upper_ba ? pg : pr, f_ba1 = upper_ba ? pr : pg;

			int cx, cy, cz, cw;
			if (mid_rg)
			{
				cx = f_rg0 * (t01.x - t11.x) + f_rg1 * (t10.x - t11.x) + ((tb.x - t11.x) << 6) + 0x80;
				cy = f_rg0 * (t01.y - t11.y) + f_rg1 * (t10.y - t11.y) + ((tb.y - t11.y) << 6) + 0x80;
			}
			else
			{
				const int bx = (upper_rg && yuv) ? t11.x : tb.x;
				const int by = (upper_rg && yuv) ? t11.y : tb.y;
				cx = f_rg0 * (t10.x - bx) + f_rg1 * (t01.x - bx) + 0x80;
				cy = f_rg0 * (t10.y - by) + f_rg1 * (t01.y - by) + 0x80;
			}
			if (mid_ba)
			{
				cz = f_ba0 * (t01.z - t11.z) + f_ba1 * (t10.z - t11.z) + ((tb.z - t11.z) << 6) + 0x80;
				cw = f_ba0 * (t01.w - t11.w) + f_ba1 * (t10.w - t11.w) + ((tb.w - t11.w) << 6) + 0x80;
			}
			else
			{
				const int bz = (upper_ba && yuv) ? t11.z : tb.z;
				const int bw = (upper_ba && yuv) ? t11.w : tb.w;
				cz = f_ba0 * (t10.z - bz) + f_ba1 * (t01.z - bz) + 0x80;
				cw = f_ba0 * (t10.w - bw) + f_ba1 * (t01.w - bw) + 0x80;
			}
			acc = vec4((cx >> 8) + pb, (cy >> 8) + pb, (cz >> 8) + pb, (cw >> 8) + pb);
		}
		else
			acc = splat4(pb);
	}
	else if (yuv)
	{
		if (sample_quad)
		{
			if (bilerp)
			{
				const bool mid_chroma = mid_texel_state && chroma_frac == 0x10 && frac_t == 0x10;
				if (mid_chroma)
				{
					acc.x = (tb.x + t10.x + t11.x + t01.x + 2) >> 2;
					acc.y = (tb.y + t10.y + t11.y + t01.y + 2) >> 2;
				}
				else
				{
					// 3-tap filter with its own upper/lower triangle decision for the chroma plane.
					const bool up = chroma_frac + frac_t >= 32;
					const int f0 = up ? 32 - frac_t : chroma_frac;
					const int f1 = up ? 32 - chroma_frac : frac_t;
					acc.x = bilerp3(up ? t11.x : tb.x, t10.x, t01.x, f0, f1);
					acc.y = bilerp3(up ? t11.y : tb.y, t10.y, t01.y, f0, f1);
				}
				if (mid_texel)
				{
					acc.z = (tb.z + t10.z + t11.z + t01.z + 2) >> 2;
					acc.w = (tb.w + t10.w + t11.w + t01.w + 2) >> 2;
				}
				else
				{
					const bool up = frac_s + frac_t >= 32;
					const int f0 = up ? 32 - frac_t : frac_s;
					const int f1 = up ? 32 - frac_s : frac_t;
					acc.z = bilerp3(up ? t11.z : tb.z, t10.z, t01.z, f0, f1);
					acc.w = bilerp3(up ? t11.w : tb.w, t10.w, t01.w, f0, f1);
				}
			}
			else
			{
				const bool up_l = frac_s + frac_t >= 32;
				const bool up_c = chroma_frac + frac_t >= 32;
				acc.x = up_c ? t11.x : tb.x;
				acc.y = up_c ? t11.y : tb.y;
				acc.z = up_l ? t11.z : tb.z;
				acc.w = up_l ? t11.w : tb.w;
			}
		}
		else
			acc = tb;
	}
	else if (mid_texel)
	{
		acc = vec4((tb.x + t01.x + t10.x + t11.x + 2) >> 2, (tb.y + t01.y + t10.y + t11.y + 2) >> 2,
		           (tb.z + t01.z + t10.z + t11.z + 2) >> 2, (tb.w + t01.w + t10.w + t11.w + 2) >> 2);
	}
	else if (bilerp && (sample_quad || tlut))
	{
		// The base texel already is t11 for the upper triangle; fractions flip accordingly.
		const int f0 = sum_frac >= 32 ? 32 - frac_t : frac_s;
		const int f1 = sum_frac >= 32 ? 32 - frac_s : frac_t;
		acc = vec4(bilerp3(tb.x, t10.x, t01.x, f0, f1), bilerp3(tb.y, t10.y, t01.y, f0, f1),
		           bilerp3(tb.z, t10.z, t01.z, f0, f1), bilerp3(tb.w, t10.w, t01.w, f0, f1));
	}
	else
		acc = tb;

	if (!bilerp && !convert_one)
		acc = texture_convert(acc, k);
	return acc;
}

// LOD / tile selection for 2-cycle-style LOD: texture.h:925-1003
RDP_HD void compute_lod(uint32_t &tile0, uint32_t &tile1, int &lod_frac, uint32_t max_level, int min_lod, int s, int t,
                        int sdx, int tdx, int sdy, int tdy, bool perspective_overflow, uint32_t rf)
{
	const bool tex_lod_en = (rf & RF_TEX_LOD) != 0;
	const bool sharpen = (rf & RF_SHARPEN) != 0;
	const bool detail = (rf & RF_DETAIL) != 0;
	bool magnify = false, distant = false;
	uint32_t tile_offset = 0;

	if (perspective_overflow)
	{
		distant = true;
		lod_frac = 0xff;
	}
	else
	{
		int dxs = sdx - s, dxt = tdx - t, dys = sdy - s, dyt = tdy - t;
		dxs ^= dxs >> 31;
		dxt ^= dxt >> 31;
		dys ^= dys >> 31;
		dyt ^= dyt >> 31;
		const int max_d = imax(imax(dxs, dys), imax(dxt, dyt));

		if (max_d >= 0x4000)
		{
			distant = true;
			lod_frac = 0xff;
			tile_offset = max_level;
		}
		else if (max_d < 32)
		{
			distant = max_level == 0u;
			magnify = true;
			if (!sharpen && !detail)
				lod_frac = distant ? 0xff : 0;
			else
				lod_frac = (imax(min_lod, max_d) << 3) + (sharpen ? -0x100 : 0);
		}
		else
		{
			const int mip_base = imax(find_msb(uint32_t(max_d >> 5)), 0);
			distant = uint32_t(mip_base) >= max_level;
			if (distant && !sharpen && !detail)
				lod_frac = 0xff;
			else
			{
				lod_frac = ((max_d << 3) >> mip_base) & 0xff;
				tile_offset = uint32_t(mip_base);
			}
		}
	}

	if (tex_lod_en)
	{
		if (distant)
			tile_offset = max_level;
		if (!detail)
		{
			tile0 = (tile0 + tile_offset) & 7u;
			if (distant || (!sharpen && magnify))
				tile1 = tile0;
			else
				tile1 = (tile0 + 1) & 7u;
		}
		else
		{
			tile1 = (tile0 + tile_offset + ((distant || magnify) ? 1u : 2u)) & 7u;
			tile0 = (tile0 + tile_offset + (magnify ? 0u : 1u)) & 7u;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Copy pipe texel fetch (64 bits / 4 pixels per clock in hardware): texture.h:361-472
// ---------------------------------------------------------------------------------------------
RDP_HD int copy_fetch_word(const TileDesc &tile, const uint8_t *tmem, int s, int t, int s_offset, bool tlut)
{
	const bool high_word = s_offset < 2;
	const bool replicate_8bpp = high_word && tile.size != 2 && !tlut;
	const int s_shamt = imin(int(tile.size), 2);
	const bool large_texel = tile.size == 3;
	const uint32_t idx_mask = (large_texel || tlut) ? 0x3ffu : 0x7ffu;
	const bool mirror_s = (tile.flags & TF_MIRROR_S) != 0;
	const bool mirror_t = (tile.flags & TF_MIRROR_T) != 0;
	int samp;

	if (replicate_8bpp)
	{
		s += 2 * s_offset;
		const int sa = mask_coord(s, tile.mask_s, mirror_s);
		const int sb = mask_coord(s + 1, tile.mask_s, mirror_s);
		const int tt = mask_coord(t, tile.mask_t, mirror_t);
		const uint32_t tbase = tile.offset + tile.stride * uint32_t(tt);
		uint32_t na = (tbase * 2u + (uint32_t(sa) << s_shamt)) & 0x1fffu;
		uint32_t nb = (tbase * 2u + (uint32_t(sb) << s_shamt)) & 0x1fffu;
		na ^= (uint32_t(tt) & 1u) * 8u;
		nb ^= (uint32_t(tt) & 1u) * 8u;
		const uint32_t ia = (na >> 2) & idx_mask, ib = (nb >> 2) & idx_mask;
		int a = int(tmem_read16(tmem, ia ^ 1u));
		int b = int(tmem_read16(tmem, ib ^ 1u));
		if (tile.size == 1)
		{
			a = (a >> (8 - 4 * int(na & 2u))) & 0xff;
			b = (b >> (8 - 4 * int(nb & 2u))) & 0xff;
		}
		else if (tile.size == 0)
		{
			a = ((a >> (12 - 4 * int(na & 3u))) & 0xf) * 0x11;
			b = ((b >> (12 - 4 * int(nb & 3u))) & 0xf) * 0x11;
		}
		else
		{
			a >>= 8;
			b >>= 8;
		}
		samp = (a << 8) | b;
	}
	else
	{
		s += s_offset;
		const int ss = mask_coord(s, tile.mask_s, mirror_s);
		const int tt = mask_coord(t, tile.mask_t, mirror_t);
		const uint32_t tbase = tile.offset + tile.stride * uint32_t(tt);
		uint32_t n = (tbase * 2u + (uint32_t(ss) << s_shamt)) & 0x1fffu;
		n ^= (uint32_t(tt) & 1u) * 8u;
		const uint32_t idx = (n >> 2) & idx_mask;
		samp = int(tmem_read16(tmem, idx ^ 1u));
		if (tlut)
		{
			if (tile.size == 0)
			{
				samp = (samp >> (12 - 4 * int(n & 3u))) & 0xf;
				samp |= int(tile.palette) << 4;
			}
			else
				samp = (samp >> (8 - 4 * int(n & 2u))) & 0xff;
			samp = (samp << 2) + s_offset;
			samp = int(tmem_read16(tmem, uint32_t(samp | 0x400) ^ 1u));
		}
	}
	return samp;
}

RDP_HD int copy_fetch(const TileDesc &tile, const uint8_t *tmem, int s, int t, int s_offset, bool tlut, uint32_t fb_fmt)
{
	s = shift_texcoord(s, tile.shift_s) - (int(tile.slo) << 3);
	t = shift_texcoord(t, tile.shift_t) - (int(tile.tlo) << 3);
	s >>= 5;
	t >>= 5;
	if (fb_fmt == FB_I4)
		return 0;
	if (fb_fmt == FB_I8)
	{
		int samp = copy_fetch_word(tile, tmem, s, t, s_offset >> 1, tlut);
		return (samp >> (8 - 8 * (s_offset & 1))) & 0xff;
	}
	return copy_fetch_word(tile, tmem, s, t, s_offset, tlut);
}

// ---------------------------------------------------------------------------------------------
// Colour combiner: combiner.h:28-282.  Selector numbering is the hardware's (rdp_common.hpp:90-180);
// constant inputs are read from the state snapshot directly.
// ---------------------------------------------------------------------------------------------
struct CombinerRegs
{
	Vec4 combined, texel0, texel1, shade;
	int lod_frac, noise;
};

RDP_HD void rgb_of_color(const uint8_t *c, int &r, int &g, int &b)
{
	r = c[0];
	g = c[1];
	b = c[2];
}

RDP_HD void select_rgb_common(const CombinerRegs &in, const PassState &st, int sel, int &r, int &g, int &b, bool &hit)
{
	hit = true;
	switch (sel)
	{
	case 0: r = in.combined.x; g = in.combined.y; b = in.combined.z; break;
	case 1: r = in.texel0.x; g = in.texel0.y; b = in.texel0.z; break;
	case 2: r = in.texel1.x; g = in.texel1.y; b = in.texel1.z; break;
	case 3: rgb_of_color(st.prim_color, r, g, b); break;
	case 4: r = in.shade.x; g = in.shade.y; b = in.shade.z; break;
	case 5: rgb_of_color(st.env_color, r, g, b); break;
	default: hit = false; r = g = b = 0; break;
	}
}

RDP_HD int select_alpha_abd(const CombinerRegs &in, const PassState &st, int sel)
{
	switch (sel)
	{
	case 0: return in.combined.w;
	case 1: return in.texel0.w;
	case 2: return in.texel1.w;
	case 3: return st.prim_color[3];
	case 4: return in.shade.w;
	case 5: return st.env_color[3];
	case 6: return 0x100;
	default: return 0;
	}
}

RDP_HD int combine_channel(int a, int b, int c, int d)
{
	c = sign_extend<9>(uint32_t(c));
	a = special_9bit_expand(a);
	b = special_9bit_expand(b);
	d = special_9bit_expand(d);
	return (((a - b) * c + 0x80) >> 8) + d;
}

template <class M>
RDP_HD Vec4 run_combiner_equation(const CombinerRegs &in, const PassState &st, const M &m, int cycle)
{
	const int rs[4] = { m.comb_rgb(cycle, 0), m.comb_rgb(cycle, 1), m.comb_rgb(cycle, 2), m.comb_rgb(cycle, 3) };
	const int as[4] = { m.comb_alpha(cycle, 0), m.comb_alpha(cycle, 1), m.comb_alpha(cycle, 2), m.comb_alpha(cycle, 3) };
	int ar, ag, ab, br, bg, bb, cr, cg, cb, dr, dg, db;
	bool hit;

	select_rgb_common(in, st, rs[0], ar, ag, ab, hit);
	if (!hit)
	{
		if (rs[0] == 6) ar = ag = ab = 0x100;
		else if (rs[0] == 7) ar = ag = ab = in.noise;
	}

	select_rgb_common(in, st, rs[1], br, bg, bb, hit);
	if (!hit)
	{
		if (rs[1] == 6) rgb_of_color(st.key_center, br, bg, bb);
		else if (rs[1] == 7) br = bg = bb = st.convert[4] & 0x1ff;
	}

	select_rgb_common(in, st, rs[2], cr, cg, cb, hit);
	if (!hit)
	{
		switch (rs[2])
		{
		case 6: rgb_of_color(st.key_scale, cr, cg, cb); break;
		case 7: cr = cg = cb = in.combined.w; break;
		case 8: cr = cg = cb = in.texel0.w; break;
		case 9: cr = cg = cb = in.texel1.w; break;
		case 10: cr = cg = cb = st.prim_color[3]; break;
		case 11: cr = cg = cb = in.shade.w; break;
		case 12: cr = cg = cb = st.env_color[3]; break;
		case 13: cr = cg = cb = in.lod_frac; break;
		case 14: cr = cg = cb = st.prim_lod_frac; break;
		case 15: cr = cg = cb = st.convert[5] & 0x1ff; break;
		default: break;
		}
	}

	select_rgb_common(in, st, rs[3], dr, dg, db, hit);
	if (!hit && rs[3] == 6)
		dr = dg = db = 0x100;

	const int aa = select_alpha_abd(in, st, as[0]);
	const int ba = select_alpha_abd(in, st, as[1]);
	const int da = select_alpha_abd(in, st, as[3]);
	int ca;
	switch (as[2])
	{
	case 0: ca = in.lod_frac; break;
	case 1: ca = in.texel0.w; break;
	case 2: ca = in.texel1.w; break;
	case 3: ca = st.prim_color[3]; break;
	case 4: ca = in.shade.w; break;
	case 5: ca = st.env_color[3]; break;
	case 6: ca = st.prim_lod_frac; break;
	default: ca = 0; break;
	}

	return vec4(combine_channel(ar, br, cr, dr), combine_channel(ag, bg, cg, dg), combine_channel(ab, bb, cb, db),
	            combine_channel(aa, ba, ca, da));
}

// First cycle of 2-cycle mode: result stays unclamped; yields the alpha-test reference.
template <class M>
RDP_HD Vec4 combiner_first_cycle(const CombinerRegs &in, const PassState &st, const M &m, int alpha_dith,
                                 int coverage, uint32_t rf, int &alpha_ref)
{
	Vec4 c = run_combiner_equation(in, st, m, 0);
	if (rf & RF_ALPHA_TEST)
	{
		const int clamped = clamp_9bit(c.w);
		int expanded = clamped + ((clamped + 1) >> 8);
		if (rf & RF_ALPHA_CVG_SELECT)
			expanded = (rf & RF_CVG_TIMES_ALPHA) ? ((expanded * coverage + 4) >> 3) : (coverage << 5);
		else
			expanded += alpha_dith;
		alpha_ref = iclamp(expanded, 0, 0xff);
	}
	else
		alpha_ref = 0;
	return c;
}

// Last (or only) cycle: clamps, and may fold coverage into alpha.
template <class M>
RDP_HD Vec4 combiner_last_cycle(const CombinerRegs &in, const PassState &st, const M &m, int alpha_dith,
                                int &coverage, uint32_t rf)
{
	Vec4 c = run_combiner_equation(in, st, m, 1);
	c.x = clamp_9bit(c.x);
	c.y = clamp_9bit(c.y);
	c.z = clamp_9bit(c.z);
	c.w = clamp_9bit(c.w);

	int expanded = c.w + ((c.w + 1) >> 8);
	int modulated;
	if (rf & RF_CVG_TIMES_ALPHA)
	{
		modulated = (expanded * coverage + 4) >> 3;
		coverage = modulated >> 5;
	}
	else
		modulated = coverage << 5;

	if (rf & RF_ALPHA_CVG_SELECT)
		expanded = modulated;
	else
		expanded += alpha_dith;
	c.w = iclamp(expanded, 0, 0xff);
	return c;
}

// ---------------------------------------------------------------------------------------------
// Dither coefficients: dither.h:26-68
// ---------------------------------------------------------------------------------------------
RDP_HD int dither_matrix(int which, int idx)
{
	// magic = { 0, 6, 1, 7, 4, 2, 5, 3, 3, 5, 2, 4, 7, 1, 6, 0 }, bayer = { 0, 4, 1, 5, 4, 0, 5, 1, 3, 7, 2, 6, 7, 3, 6, 2 },
	// one nibble per entry so the look-up is a shift instead of a table in local memory.
	const uint64_t magic = 0x0617425335247160ull;
	const uint64_t bayer = 0x2637627315045140ull;
	return int(((which ? bayer : magic) >> (idx * 4)) & 7u);
}

RDP_HD void dither_coefficients(int x, int y, int mode_rgb, int mode_alpha, uint32_t noise, int &rgb_dither,
                                int &alpha_dither)
{
	const int idx = (y & 3) * 4 + (x & 3);
	if (mode_rgb < 2)
		rgb_dither = dither_matrix(mode_rgb, idx) * 0x49;
	else if (mode_rgb == 2)
		rgb_dither = int(noise & 0x1ffu);
	else
		rgb_dither = 0;

	if (mode_alpha == 3)
		alpha_dither = 0;
	else if (mode_alpha == 2)
		alpha_dither = int(noise & 7u);
	else
	{
		alpha_dither = mode_rgb >= 2 ? dither_matrix(mode_rgb & 1, idx) : (rgb_dither & 7);
		if (mode_alpha == 1)
			alpha_dither = ~alpha_dither & 7;
	}
}

// ---------------------------------------------------------------------------------------------
// shade_pixel: shading.h:38-359
// ---------------------------------------------------------------------------------------------
// The caller has already resolved the primitive, its span for line y (valid) and its state; fill-mode states never
// get here (process_pixel_primitive handles them).  `m` holds the state's static modes, literal under a Spec.
template <class S>
RDP_HD bool depth_early_reject(int z, const PrimSetup &p, const EarlyDepth &e, int coverage_count, uint32_t bf, int z_mode);

// Colour half of shade_pixel: everything after coverage and depth -- shade interpolation, texture coordinates,
// LOD, TMEM sampling, both combiner cycles, dither coefficients.  `coverage` is the 8-bit sub-sample mask of the
// pixel (non-zero); `coverage_count` enters as its population count and leaves modulated by alpha when the state
// asks for it.  Shared by the in-order path (shade_pixel) and the deferred tile kernel's compacted shade phase.
struct ShadedColor
{
	Vec4 combined;   // combiner output, alpha already merged with coverage / dither
	int alpha_ref;   // what the alpha test compares
	int rgb_dither;  // 9-bit rgb dither pattern
	int shade_alpha; // shade alpha + alpha dither (blender input 1B)
	uint32_t noise;  // noise word the alpha-test dither reads
};

template <class S>
RDP_HD void shade_color(const ShadeContext &ctx, int x, int y, uint32_t prim_index, const PrimSetup &p, const Span &sp,
                        const PassState &st, const StaticModes<S> &m, int coverage, int &coverage_count, ShadedColor &out)
{
	const uint32_t rf = m.rf;
	const uint8_t *tmem = ctx.tmem + size_t(p.tmem_instance) * kTmemBytes;
	const bool flip = (p.flags & SF_FLIP) != 0;
	const bool perspective = (rf & RF_PERSPECTIVE) != 0;

	uint32_t noise = 0;
	if (rf & RF_NOISE)
		noise = noise_hash(uint32_t(x), uint32_t(y), prim_index + ctx.base_prim_index);

	const int sl = ctx.scale_log2;
	const int dx = x - sp.base_x;
	const int direction = flip ? 1 : -1;
	const bool uses_lod = (rf & RF_USES_LOD) != 0;

	CombinerRegs regs;
	regs.shade = interpolate_shade(sp, p, dx, coverage, sl);
	regs.combined = splat4(0);
	regs.texel0 = splat4(0);
	regs.texel1 = splat4(0);
	regs.lod_frac = 0;
	regs.noise = int(((noise & 7u) << 6) | 0x20u);

	// Texture coordinates (and the neighbours LOD needs).
	int s, t;
	int s_dx = 0, t_dx = 0, s_dy = 0, t_dy = 0;
	bool persp_overflow = false;
	{
		const int ms = (p.dstzw_dx[0] & ~0x1f) >> sl, mt = (p.dstzw_dx[1] & ~0x1f) >> sl, mw = (p.dstzw_dx[3] & ~0x1f) >> sl;
		const int rs = int(uint32_t(sp.stzw[0]) + uint32_t(ms) * uint32_t(dx));
		const int rt = int(uint32_t(sp.stzw[1]) + uint32_t(mt) * uint32_t(dx));
		const int rw = int(uint32_t(sp.stzw[3]) + uint32_t(mw) * uint32_t(dx));
		divide_st(perspective, rs >> 16, rt >> 16, rw >> 16, s, t, persp_overflow);
		if (uses_lod)
		{
			// native_texture_lod quirk: LOD as the native resolution would compute it (shading.h:193-196)
			const int lod_dir = (sl != 0 && (st.quirks & QUIRK_NATIVE_TEXTURE_LOD)) ? (direction << sl) : direction;
			const int xs = int(uint32_t(rs) + uint32_t(lod_dir * ms));
			const int xt = int(uint32_t(rt) + uint32_t(lod_dir * mt));
			const int xw = int(uint32_t(rw) + uint32_t(lod_dir * mw));
			const int dy_mul = lod_dir < 0 ? -lod_dir : lod_dir; // |direction|: 1, or S under the native-LOD quirk (interpolation.h:104-107)
			const int ys = int(uint32_t(rs) + uint32_t(dy_mul) * uint32_t((p.dstzw_dy[0] & ~0x7fff) >> sl));
			const int yt = int(uint32_t(rt) + uint32_t(dy_mul) * uint32_t((p.dstzw_dy[1] & ~0x7fff) >> sl));
			const int yw = int(uint32_t(rw) + uint32_t(dy_mul) * uint32_t((p.dstzw_dy[3] & ~0x7fff) >> sl));
			divide_st(perspective, xs >> 16, xt >> 16, xw >> 16, s_dx, t_dx, persp_overflow);
			divide_st(perspective, ys >> 16, yt >> 16, yw >> 16, s_dy, t_dy, persp_overflow);
		}
	}
	uint32_t tile0 = uint32_t(p.tile) & 7u;
	uint32_t tile1 = (tile0 + 1u) & 7u;
	const uint32_t max_level = uint32_t(p.tile) >> 3;
	if (uses_lod)
		compute_lod(tile0, tile1, regs.lod_frac, max_level, st.min_lod, s, t, s_dx, t_dx, s_dy, t_dy, persp_overflow, rf);

	const bool convert_one = (rf & RF_CONVERT_ONE) != 0;
	const bool bilerp0 = (rf & RF_BILERP0) != 0;
	const bool bilerp1 = (rf & RF_BILERP1) != 0;
	bool uses_texel1 = (rf & RF_USES_TEXEL1) != 0;

	if (rf & RF_USES_TEXEL0)
		regs.texel0 = sample_texture<S::tk_mask>(st.tiles[tile0], tmem, s, t, rf, bilerp0, false, st.convert, splat4(0));

	if (rf & RF_USES_PIPELINED_TEXEL1)
	{
		// 1-cycle mode reading TEXEL1 observes the next pixel's texel (or the next line's first pixel).
		const Span &next = ctx.spans[(p.span_base << sl) + uint32_t(y - (int(p.line_lo) << sl) + 1)];
		const bool long_span = sp.lodlength >= 8;
		const bool end_span = x == (flip ? sp.end_x : sp.start_x);
		bool dummy = false;
		if (end_span && long_span && next.valid)
			divide_st(perspective, next.stzw[0] >> 16, next.stzw[1] >> 16, next.stzw[3] >> 16, s, t, dummy);
		else
		{
			const int ndx = dx + (direction << sl);
			const int rs = int(uint32_t(sp.stzw[0]) + uint32_t((p.dstzw_dx[0] & ~0x1f) >> sl) * uint32_t(ndx));
			const int rt = int(uint32_t(sp.stzw[1]) + uint32_t((p.dstzw_dx[1] & ~0x1f) >> sl) * uint32_t(ndx));
			const int rw = int(uint32_t(sp.stzw[3]) + uint32_t((p.dstzw_dx[3] & ~0x1f) >> sl) * uint32_t(ndx));
			divide_st(perspective, rs >> 16, rt >> 16, rw >> 16, s, t, dummy);
		}
		tile1 = tile0;
		uses_texel1 = true;
	}

	if (uses_texel1)
	{
		if (convert_one && !bilerp1)
			regs.texel1 = texture_convert(regs.texel0, st.convert);
		else
			regs.texel1 = sample_texture<S::tk_mask>(st.tiles[tile1], tmem, s, t, rf, bilerp1, convert_one, st.convert, regs.texel0);
	}

	int alpha_dith;
	dither_coefficients(x, (rf & RF_INTERLACE) ? (y >> 1) : y, m.dither_rgb(), m.dither_alpha(), noise, out.rgb_dither, alpha_dith);

	if (rf & RF_TWO_CYCLE)
	{
		regs.combined = combiner_first_cycle(regs, st, m, alpha_dith, coverage_count, rf, out.alpha_ref);
		// Pipelining: TEXEL1 is seen as TEXEL0 by the second cycle and vice versa.
		const Vec4 tmp = regs.texel0;
		regs.texel0 = regs.texel1;
		regs.texel1 = tmp;
		if (rf & RF_NOISE_DUAL)
		{
			const uint32_t n2 = noise_hash(uint32_t(x + 1023), uint32_t(y + 7), prim_index + ctx.base_prim_index + 11u);
			regs.noise = int(((n2 & 7u) << 6) | 0x20u);
			// NB: the reseed also changes the noise seen by the alpha-test dither below (noise.h keeps one seed).
			noise = n2;
		}
		out.combined = combiner_last_cycle(regs, st, m, alpha_dith, coverage_count, rf);
	}
	else
	{
		out.combined = combiner_last_cycle(regs, st, m, alpha_dith, coverage_count, rf);
		out.alpha_ref = out.combined.w;
	}
	out.shade_alpha = imin(regs.shade.w + alpha_dith, 0xff);
	out.noise = noise;
}

// Coverage half: the 8-bit sub-sample mask of pixel x on this span, 0 when the primitive does not shade it
// (no sample covered, or -- without AA -- the first sample not covered).
RDP_HD int shade_coverage(const Span &sp, int x, uint32_t rf, int scale_log2 = 0)
{
	const int coverage = compute_coverage(sp, x, scale_log2 > 2);
	if (!(rf & RF_AA) && (coverage & 1) == 0)
		return 0;
	return coverage;
}

// Returns 0: not shaded (no coverage / alpha test), 1: shaded, `out` valid, 2: rejected by the early depth test.
template <class S>
RDP_HD int shade_pixel(const ShadeContext &ctx, int x, int y, uint32_t prim_index, const PrimSetup &p, const Span &sp,
                       const PassState &st, const StaticModes<S> &m, const EarlyDepth &early, Shaded &out)
{
	const uint32_t rf = m.rf;

	if (rf & RF_COPY)
	{
		const uint8_t *tmem = ctx.tmem + size_t(p.tmem_instance) * kTmemBytes;
		const bool flip = (p.flags & SF_FLIP) != 0;
		const bool perspective = (rf & RF_PERSPECTIVE) != 0;
		if (x < sp.start_x || x > sp.end_x)
			return 0;

		int dx = flip ? (x - sp.start_x) : (sp.end_x - x);
		dx >>= ctx.scale_log2; // the copy pipe duplicates pixels when upscaling (interpolation.h:60-61)
		int dx_mask, dx_shift;
		switch (ctx.fb_fmt)
		{
		case FB_I4: dx_mask = 0; dx_shift = 0; break;
		case FB_I8: dx_mask = ~7; dx_shift = 3; break;
		case FB_RGBA8888: dx_mask = 0; dx_shift = 1; break; // copy into 32bpp is undefined on hardware
		default: dx_mask = ~3; dx_shift = 2; break;
		}
		const int snapped_dx = dx & dx_mask;
		const int s_offset = dx - snapped_dx;
		const int lerp_dx = (dx >> dx_shift) * (flip ? 1 : -1);
		const int s = int(uint32_t(sp.stzw[0]) + uint32_t(p.dstzw_dx[0] & ~0x1f) * uint32_t(lerp_dx)) >> 16;
		const int t = int(uint32_t(sp.stzw[1]) + uint32_t(p.dstzw_dx[1] & ~0x1f) * uint32_t(lerp_dx)) >> 16;
		const int w = int(uint32_t(sp.stzw[3]) + uint32_t(p.dstzw_dx[3] & ~0x1f) * uint32_t(lerp_dx)) >> 16;
		int ps, pt;
		bool ovf = false;
		divide_st(perspective, s, t, w, ps, pt, ovf);

		const TileDesc &tile = st.tiles[p.tile & 7u];
		const int texel = copy_fetch(tile, tmem, ps, pt, s_offset, (rf & RF_TLUT) != 0, ctx.fb_fmt);
		out.raw = uint32_t(texel);
		out.kind = 2;
		const uint32_t fb_bytes = ctx.fb_fmt == FB_I4 ? 0u : (ctx.fb_fmt == FB_I8 ? 1u : (ctx.fb_fmt == FB_RGBA8888 ? 4u : 2u));
		if ((rf & RF_ALPHA_TEST) && fb_bytes == 2 && (texel & 1) == 0)
			return 0;
		return 1;
	}

	const int coverage = shade_coverage(sp, x, rf, ctx.scale_log2);
	if (coverage == 0)
		return 0;
	int coverage_count = pop_count(uint32_t(coverage));

	const int z = interpolate_depth(sp, p, x - sp.base_x, coverage, ctx.scale_log2);
	if ((m.bf & BF_Z_COMPARE) && !(rf & RF_CVG_TIMES_ALPHA) && depth_early_reject(z, p, early, coverage_count, m.bf, m.z_mode()))
		return 2;

	ShadedColor c;
	shade_color<S>(ctx, x, y, prim_index, p, sp, st, m, coverage, coverage_count, c);

	if ((rf & RF_AA) && coverage_count == 0)
		return 0;

	if (rf & RF_ALPHA_TEST)
	{
		const int threshold = (rf & RF_ALPHA_TEST_DITHER) ? int(c.noise & 0xffu) : int(st.blend_color[3]);
		if (c.alpha_ref < threshold)
			return 0;
	}

	out.r = c.combined.x & 0xff;
	out.g = c.combined.y & 0xff;
	out.b = c.combined.z & 0xff;
	out.a = c.combined.w & 0xff;
	out.z = z;
	out.dither = c.rgb_dither;
	out.coverage = coverage_count;
	out.shade_alpha = c.shade_alpha;
	out.kind = 0;
	return 1;
}
}
