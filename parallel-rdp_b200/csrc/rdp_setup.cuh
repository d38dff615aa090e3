// rdp_setup.cuh -- primitive decode and span (edge-walk) set-up, callable from host and device.
//
// Behaviour follows the reference bit for bit; structure is ours:
//   * command word decode: parallel-rdp/rdp_device.cpp:268-344 (triangles), :713-821 (rectangles)
//   * per-primitive fix-ups and derived depth delta: rdp_renderer.cpp:717-741, :910-925, :1402-1423, :1489-1503
//   * span set-up: shaders/span_setup.comp:61-227
// In the reference the first two run on the CPU per primitive; here they run in a GPU kernel so the host
// only forwards raw command words.
#pragma once
#include "rdp_types.h"

namespace rdpb
{
RDP_HD int find_msb(uint32_t v)
{
#if defined(__CUDA_ARCH__)
	return 31 - __clz(int(v));
#else
	return v ? 31 - __builtin_clz(v) : -1;
#endif
}

RDP_HD int find_lsb(uint32_t v)
{
#if defined(__CUDA_ARCH__)
	return __ffs(int(v)) - 1;
#else
	return v ? __builtin_ctz(v) : -1;
#endif
}

RDP_HD int pop_count(uint32_t v)
{
#if defined(__CUDA_ARCH__)
	return __popc(v);
#else
	return __builtin_popcount(v);
#endif
}

RDP_HD int imin(int a, int b) { return a < b ? a : b; }
RDP_HD int imax(int a, int b) { return a > b ? a : b; }
RDP_HD int iclamp(int v, int lo, int hi) { return imin(imax(v, lo), hi); }

template <int Bits>
RDP_HD int32_t sign_extend(uint32_t v)
{
	return int32_t(v << (32 - Bits)) >> (32 - Bits);
}

// Number of 32-bit words of a command, by opcode (hardware table; cf. angrylion rdp.c:391-462).
RDP_HD uint32_t command_length_words(uint32_t op)
{
	switch (op)
	{
	case OP_TRI_FILL: return 8;
	case OP_TRI_FILL_Z: return 12;
	case OP_TRI_TEX: return 24;
	case OP_TRI_TEX_Z: return 28;
	case OP_TRI_SHADE: return 24;
	case OP_TRI_SHADE_Z: return 28;
	case OP_TRI_SHADE_TEX: return 40;
	case OP_TRI_SHADE_TEX_Z: return 44;
	case OP_TEX_RECT:
	case OP_TEX_RECT_FLIP: return 4;
	default: return 2;
	}
}

RDP_HD bool op_is_primitive(uint32_t op)
{
	return (op >= OP_TRI_FILL && op <= OP_TRI_SHADE_TEX_Z) || op == OP_TEX_RECT || op == OP_TEX_RECT_FLIP ||
	       op == OP_FILL_RECT;
}

// Coefficient blocks interleave the integer halves (first 4 words) and the fractional halves (next 4).
RDP_HD int32_t coeff_hi(const uint32_t *w, int int_word, int frac_word)
{
	return int32_t((w[int_word] & 0xffff0000u) | (w[frac_word] >> 16));
}
RDP_HD int32_t coeff_lo(const uint32_t *w, int int_word, int frac_word)
{
	return int32_t((w[int_word] << 16) | (w[frac_word] & 0xffffu));
}

// 16 words -> value, d/dx, d/de, d/dy of four channels.  With three_only the 4th channel of each pair is skipped
// (texture coefficients carry S,T,W and an unused slot).
RDP_HD void decode_coeff_block(const uint32_t *w, int32_t *v, int32_t *dx, int32_t *de, int32_t *dy, bool three_only,
                               int c0, int c1, int c2, int c3)
{
	v[c0] = coeff_hi(w, 0, 4);
	v[c1] = coeff_lo(w, 0, 4);
	v[c2] = coeff_hi(w, 1, 5);
	dx[c0] = coeff_hi(w, 2, 6);
	dx[c1] = coeff_lo(w, 2, 6);
	dx[c2] = coeff_hi(w, 3, 7);
	de[c0] = coeff_hi(w, 8, 12);
	de[c1] = coeff_lo(w, 8, 12);
	de[c2] = coeff_hi(w, 9, 13);
	dy[c0] = coeff_hi(w, 10, 14);
	dy[c1] = coeff_lo(w, 10, 14);
	dy[c2] = coeff_hi(w, 11, 15);
	if (!three_only)
	{
		v[c3] = coeff_lo(w, 1, 5);
		dx[c3] = coeff_lo(w, 3, 7);
		de[c3] = coeff_lo(w, 9, 13);
		dy[c3] = coeff_lo(w, 11, 15);
	}
}

// The 4-bit "log2-ish" encoding of a depth delta as written to the Z-buffer's dz field.
RDP_HD uint32_t dz_compress_bits(uint32_t dz)
{
	uint32_t v = 0;
	if (dz & 0xff00u) v |= 8;
	if (dz & 0xf0f0u) v |= 4;
	if (dz & 0xccccu) v |= 2;
	if (dz & 0xaaaau) v |= 1;
	return v;
}

RDP_HD int normalize_dz(int dz)
{
	if (dz >= 0x8000)
		return 0x8000;
	if (dz == 0)
		return 1;
	return 1 << (find_msb(uint32_t(dz)) + 1);
}

// Decode one primitive command into a PrimSetup.  `w` points at the first command word.
RDP_HD void decode_primitive(const uint32_t *w, const PassState &st, const PrimRecord &rec, PrimSetup &p)
{
	const uint32_t op = (w[0] >> 24) & 63u;
	const bool fill_or_copy = (st.raster_flags & (RF_FILL | RF_COPY)) != 0;
	const bool copy = (st.raster_flags & RF_COPY) != 0;

	uint32_t flags = 0;
	for (int i = 0; i < 4; i++)
	{
		p.rgba[i] = p.drgba_dx[i] = p.drgba_de[i] = p.drgba_dy[i] = 0;
		p.stzw[i] = p.dstzw_dx[i] = p.dstzw_de[i] = p.dstzw_dy[i] = 0;
	}

	if (op == OP_FILL_RECT || op == OP_TEX_RECT || op == OP_TEX_RECT_FLIP)
	{
		uint32_t xl = (w[0] >> 12) & 0xfffu;
		uint32_t yl = w[0] & 0xfffu;
		uint32_t xh = (w[1] >> 12) & 0xfffu;
		uint32_t yh = w[1] & 0xfffu;
		if (fill_or_copy)
			yl |= 3;

		p.xh = int32_t(xh << 13);
		p.xl = int32_t(xl << 13);
		p.xm = int32_t(xl << 13);
		p.dxhdy = p.dxmdy = p.dxldy = 0;
		p.yh = int16_t(yh);
		p.ym = int16_t(yl);
		p.yl = int16_t(yl);
		p.tile = 0;
		flags |= SF_FLIP;

		if (op != OP_FILL_RECT)
		{
			p.tile = uint8_t((w[1] >> 24) & 7u);
			int32_t s = int32_t((w[2] >> 16) & 0xffffu);
			int32_t t = int32_t(w[2] & 0xffffu);
			int32_t dsdx = sign_extend<16>(w[3] >> 16);
			int32_t dtdy = sign_extend<16>(w[3] & 0xffffu);
			p.stzw[0] = int32_t(uint32_t(s) << 16);
			p.stzw[1] = int32_t(uint32_t(t) << 16);
			if (op == OP_TEX_RECT)
			{
				p.dstzw_dx[0] = int32_t(uint32_t(dsdx) << 11);
				p.dstzw_de[1] = int32_t(uint32_t(dtdy) << 11);
				p.dstzw_dy[1] = int32_t(uint32_t(dtdy) << 11);
			}
			else
			{
				p.dstzw_dx[1] = int32_t(uint32_t(dtdy) << 11);
				p.dstzw_de[0] = int32_t(uint32_t(dsdx) << 11);
				p.dstzw_dy[0] = int32_t(uint32_t(dsdx) << 11);
			}
		}
	}
	else
	{
		const bool flip = (w[0] & 0x800000u) != 0;
		const bool sign_dxhdy = (w[5] & 0x80000000u) != 0;
		if (flip)
			flags |= SF_FLIP;
		if (flip == sign_dxhdy)
			flags |= SF_DO_OFFSET;

		p.tile = uint8_t((w[0] >> 16) & 63u);
		p.yl = int16_t(sign_extend<14>(w[0]));
		p.ym = int16_t(sign_extend<14>(w[1] >> 16));
		p.yh = int16_t(sign_extend<14>(w[1]));

		// One fractional bit is dropped (the hardware ignores it), as in the reference.
		p.xl = sign_extend<28>(w[2]) >> 1;
		p.xh = sign_extend<28>(w[4]) >> 1;
		p.xm = sign_extend<28>(w[6]) >> 1;
		p.dxldy = sign_extend<28>(w[3] >> 2) >> 1;
		p.dxhdy = sign_extend<28>(w[5] >> 2) >> 1;
		p.dxmdy = sign_extend<28>(w[7] >> 2) >> 1;

		const uint32_t *c = w + 8;
		if (op & 4u) // shade
		{
			decode_coeff_block(c, p.rgba, p.drgba_dx, p.drgba_de, p.drgba_dy, false, 0, 1, 2, 3);
			c += 16;
		}
		if (op & 2u) // texture: S, T, W
		{
			decode_coeff_block(c, p.stzw, p.dstzw_dx, p.dstzw_de, p.dstzw_dy, true, 0, 1, 3, 2);
			c += 16;
		}
		if (op & 1u) // depth
		{
			p.stzw[2] = int32_t(c[0]);
			p.dstzw_dx[2] = int32_t(c[1]);
			p.dstzw_de[2] = int32_t(c[2]);
			p.dstzw_dy[2] = int32_t(c[3]);
		}
	}

	if (copy)
		flags |= SF_SKIP_XFRAC;
	if (fill_or_copy)
		flags |= SF_FILL_COPY_RASTER;
	if (st.raster_flags & RF_INTERLACE)
	{
		flags |= SF_INTERLACE;
		if (st.raster_flags & RF_KEEP_ODD)
			flags |= SF_KEEP_ODD;
	}

	// YM above the first rasterised sub-scanline can never be observed; push it out of range.
	if (int(p.ym) < (int(p.yh) & ~3))
		p.ym = 0x7fff;

	if (st.raster_flags & RF_PRIM_DEPTH)
	{
		p.stzw[2] = st.prim_depth;
		p.dstzw_dx[2] = p.dstzw_de[2] = p.dstzw_dy[2] = 0;
		p.dz = st.prim_dz;
		p.dz_compressed = uint8_t(dz_compress_bits(st.prim_dz));
	}
	else
	{
		int dzdx = p.dstzw_dx[2] >> 16;
		int dzdy = p.dstzw_dy[2] >> 16;
		int dzpix = (dzdx < 0 ? (~dzdx & 0x7fff) : dzdx) + (dzdy < 0 ? (~dzdy & 0x7fff) : dzdy);
		dzpix = normalize_dz(dzpix);
		p.dz = uint16_t(dzpix);
		p.dz_compressed = uint8_t(dz_compress_bits(uint32_t(dzpix)));
	}

	if (rec.line_hi < rec.line_lo)
		flags |= SF_NO_PIXELS;
	if (op == OP_FILL_RECT || op == OP_TEX_RECT_FLIP || (op == OP_TEX_RECT && (st.quirks & QUIRK_NATIVE_RESOLUTION_TEX_RECT)))
		flags |= SF_DISABLE_UPSCALING;

	p.flags = uint8_t(flags);
	p.span_base = rec.span_base;
	p.state_id = rec.state_id;
	p.tmem_instance = rec.tmem_instance;
	p.line_lo = rec.line_lo;
	p.line_hi = rec.line_hi;
	p.pad = 0;
}

// Host-side helper: active scanline range of a primitive given the Y scissor.  Mirrors how the reference
// sizes its span jobs (rdp_renderer.cpp:1523-1554) so that line_lo..line_hi+1 spans exist.
RDP_HD void primitive_line_range(const uint32_t *w, uint32_t raster_flags, uint32_t sc_ylo, uint32_t sc_yhi,
                                 int &line_lo, int &line_hi)
{
	const uint32_t op = (w[0] >> 24) & 63u;
	int yh, yl;
	if (op == OP_FILL_RECT || op == OP_TEX_RECT || op == OP_TEX_RECT_FLIP)
	{
		yl = int(w[0] & 0xfffu);
		yh = int(w[1] & 0xfffu);
		if (raster_flags & (RF_FILL | RF_COPY))
			yl |= 3;
	}
	else
	{
		yl = sign_extend<14>(w[0]);
		yh = sign_extend<14>(w[1]);
	}
	int lo_sub = imax(yh, int(sc_ylo));
	int hi_sub = imin(yl - 1, int(sc_yhi) - 1);
	line_lo = lo_sub >> 2;
	line_hi = hi_sub >> 2;
	if (line_hi - line_lo > 1023)
		line_hi = line_lo + 1023;
}

RDP_HD int quantize_edge_x(int x)
{
	// 16.16 -> x.3 with a sticky bit for the discarded fraction.
	return (x >> 12) | ((x & 0xfff) != 0 ? 1 : 0);
}

// sign-extend the low `bits` bits (run-time width: 27 + log2(scale))
RDP_HD int sign_extend_bits(uint32_t v, int bits)
{
	const int sh = 32 - bits;
	return int(v << sh) >> sh;
}

// span_setup.comp:80-85: a derivative times a distance in scaled pixels -- whole native pixels step by the full
// derivative, the sub-pixel remainder by the derivative quantised to the scale.
RDP_HD uint32_t step_snapped(int d, int dist, int sl)
{
	return uint32_t(dist >> sl) * uint32_t(d) + uint32_t(dist & ((1 << sl) - 1)) * uint32_t(d >> sl);
}

// Edge-walk one scanline of one primitive.  `sl` = log2 of the internal upscaling factor: y is then a scanline of
// the upscaled image and every X in the result is in upscaled pixels (span_setup.comp:87-227 with SCALING_LOG2).
RDP_HD void compute_span(const PrimSetup &p, const PassState &st, int y, Span &out, int sl = 0)
{
	const int S = 1 << sl;
	// A primitive that is not upscaled shows every upscaled line the span of its native line's first sub-line.
	const bool native_only = sl != 0 && (p.flags & SF_DISABLE_UPSCALING) != 0;
	if (native_only)
		y &= ~(S - 1);
	const bool flip = (p.flags & SF_FLIP) != 0;
	const bool do_offset = (p.flags & SF_DO_OFFSET) != 0;
	const bool skip_xfrac = (p.flags & SF_SKIP_XFRAC) != 0;

	// Attributes stepped down the major edge to this scanline.
	{
		const int dy = y - ((int(p.yh) >> 2) << sl);
		int xh = int(uint32_t(p.xh) * uint32_t(S) + uint32_t(dy) * (uint32_t(p.dxhdy) << 2));
		if (do_offset)
			xh = int(uint32_t(xh) + uint32_t(3 * S) * uint32_t(p.dxhdy));

		const int base_x = xh >> 15;
		const int xfrac = skip_xfrac ? 0 : ((xh >> 7) & 0xff);

		for (int i = 0; i < 4; i++)
		{
			int diff_c = 0, diff_t = 0;
			if (do_offset)
			{
				// Attributes are latched at the last sub-scanline instead of the first.
				int deh = p.drgba_de[i] & ~0x1ff;
				int dyh = p.drgba_dy[i] & ~0x1ff;
				diff_c = deh - (deh >> 2) - dyh + (dyh >> 2);
				deh = p.dstzw_de[i] & ~0x1ff;
				dyh = p.dstzw_dy[i] & ~0x1ff;
				diff_t = deh - (deh >> 2) - dyh + (dyh >> 2);
			}
			int c = int(uint32_t(p.rgba[i]) + step_snapped(p.drgba_de[i], dy, sl));
			c = int(uint32_t(c & ~0x1ff) + uint32_t(diff_c) - step_snapped((p.drgba_dx[i] >> 8) & ~1, xfrac, sl)) & ~0x3ff;
			out.rgba[i] = c;

			int t = int(uint32_t(p.stzw[i]) + step_snapped(p.dstzw_de[i], dy, sl));
			t = int(uint32_t(t & ~0x1ff) + uint32_t(diff_t) - step_snapped((p.dstzw_dx[i] >> 8) & ~1, xfrac, sl)) & ~0x3ff;
			out.stzw[i] = t;
		}
		out.base_x = base_x;
	}

	// Edges at the four sub-scanlines.
	const int yh_base = (int(p.yh) & ~3) * S;
	const int ym_base = int(p.ym) * S;
	const int ylo = imax(int(p.yh), int(st.sc_ylo)) * S;
	const int yhi = imin(int(p.yl), int(st.sc_yhi)) * S;
	const int lo_scissor = (int(st.sc_xlo) << 1) * S;
	const int hi_scissor = (int(st.sc_xhi) << 1) * S;

	int xleft[4], xright[4];
	bool invalid[4];
	bool all_invalid = true, all_over = true, all_under = true;
	int min_left = 0x7fffffff, max_right = -0x7fffffff - 1;

	for (int i = 0; i < 4; i++)
	{
		const int ys = y * 4 + i;
		const bool clip_y = ys < ylo || ys >= yhi;

		int xh = int(uint32_t(p.xh) * uint32_t(S) + uint32_t(ys - yh_base) * uint32_t(p.dxhdy));
		int xm = int(uint32_t(p.xm) * uint32_t(S) + uint32_t(ys - yh_base) * uint32_t(p.dxmdy));
		int xl = int(uint32_t(p.xl) * uint32_t(S) + uint32_t(ys - ym_base) * uint32_t(p.dxldy));
		if (ys < ym_base)
			xl = xm;

		// (27 + log2 scale)-bit wrap (the set-up already dropped one of the 28 bits for this headroom).
		xl = sign_extend_bits(uint32_t(xl), 27 + sl);
		xh = sign_extend_bits(uint32_t(xh), 27 + sl);

		const int xh_q = quantize_edge_x(xh);
		const int xl_q = quantize_edge_x(xl);
		int l = flip ? xh_q : xl_q;
		int r = flip ? xl_q : xh_q;

		bool inv = (l >> 1) > (r >> 1);

		all_over = all_over && (imin(l, r) >= hi_scissor);
		all_under = all_under && (imax(l, r) < lo_scissor);

		l = iclamp(l, lo_scissor, hi_scissor);
		r = iclamp(r, lo_scissor, hi_scissor);

		inv = inv || clip_y;
		if (inv)
		{
			l = 0xffff;
			r = 0;
		}
		invalid[i] = inv;
		all_invalid = all_invalid && inv;
		xleft[i] = l;
		xright[i] = r;
		min_left = imin(min_left, l);
		max_right = imax(max_right, r);
	}
	(void)invalid;

	const int start_x = min_left >> 3;
	const int end_x = max_right >> 3;
	for (int i = 0; i < 4; i++)
	{
		out.xl[i] = uint16_t(xleft[i]);
		out.xr[i] = uint16_t(xright[i]);
	}
	out.start_x = int16_t(start_x);
	out.end_x = int16_t(end_x);

	bool valid = !all_invalid && !all_over && !all_under;
	if (p.flags & SF_INTERLACE)
		if (((y >> sl) & 1) != ((p.flags & SF_KEEP_ODD) ? 1 : 0))
			valid = false;
	out.valid = valid ? 1 : 0;
	out.lodlength = int16_t(flip ? (end_x - out.base_x) : (out.base_x - start_x));
	int touch_lo = start_x, touch_hi = end_x;
	if (native_only)
	{
		// upscaled pixels x with (x & ~(S - 1)) in [start_x, end_x]
		touch_lo = (start_x + (S - 1)) & ~(S - 1);
		touch_hi = (end_x & ~(S - 1)) + (S - 1);
	}
	out.touch_range = uint32_t(uint16_t(int16_t(touch_lo))) | (uint32_t(uint16_t(int16_t(touch_hi))) << 16);
}
}
