// rdp_tmem.cuh -- TMEM upload (LoadTile / LoadBlock / LoadTLUT) evaluated as a gather: every one of the
// 2048 16-bit TMEM words works out, for each upload in order, whether that upload wrote it and from which
// RDRAM address.  That makes the replay embarrassingly parallel over TMEM words while uploads stay ordered.
//
// Behavioural reference (bit-exact target): parallel-rdp/shaders/tmem_update.comp:101-577; the host-side
// decomposition of the load commands into TmemUpload records follows rdp_renderer.cpp:3099-3486.
#pragma once
#include "rdp_setup.cuh"

namespace rdpb
{
enum
{
	UPLOAD_TILE = 0,
	UPLOAD_TLUT = 1,
	UPLOAD_BLOCK = 2,
};

RDP_HD int upload_row_of(int offset, float inv_stride)
{
	// Exact for all relevant inputs; deliberately evaluated as (float(offset) + 0.5f) * inv with no contraction.
#if defined(__CUDA_ARCH__)
	return int(__fmul_rn(__fadd_rn(float(offset), 0.5f), inv_stride));
#else
	volatile float biased = float(offset) + 0.5f;
	volatile float prod = biased * inv_stride;
	return int(prod);
#endif
}

struct RdramView
{
	const uint8_t *data;
	uint32_t mask; // size - 1
};

RDP_HD uint32_t rdram_read8(const RdramView &r, uint32_t addr) { return r.data[(addr & r.mask) ^ 3u]; }
RDP_HD uint32_t rdram_read16_aligned(const RdramView &r, uint32_t addr)
{
	return reinterpret_cast<const uint16_t *>(r.data)[((addr & r.mask) >> 1) ^ 1u];
}
RDP_HD uint32_t rdram_read16(const RdramView &r, uint32_t addr)
{
	if ((addr & 1u) == 0)
		return rdram_read16_aligned(r, addr);
	return (rdram_read8(r, addr) << 8) | rdram_read8(r, addr + 1);
}
RDP_HD uint32_t rdram_read32(const RdramView &r, uint32_t addr)
{
	if ((addr & 3u) == 0)
		return reinterpret_cast<const uint32_t *>(r.data)[(addr & r.mask) >> 2];
	return (rdram_read8(r, addr) << 24) | (rdram_read8(r, addr + 1) << 16) | (rdram_read8(r, addr + 2) << 8) |
	       rdram_read8(r, addr + 3);
}

// 16-bit-granular tiles (4/8/16 bpp, non-YUV).  Returns true and the new word if this upload wrote `word_index`.
RDP_HD bool tmem_gather_16(const TmemUpload &u, const RdramView &rdram, int word_index, uint32_t &value)
{
	const int base = (u.tmem_offset & 0xfff) >> 1;
	const int stride = u.tmem_stride_words;
	int pixel = (word_index - base) & 0x7ff;
	int ux, uy;
	int x_xor = 0;

	if (u.mode == UPLOAD_BLOCK)
	{
		const int qword = pixel >> 2;
		if (stride == 0)
			x_xor = (((qword * u.dxt) >> 16) & 1) << 1;
		else
		{
			// T is floor(x * dxt / 2048); solve qword = x + T(x) * stride for x by testing the few candidate T.
			int min_t = upload_row_of(qword, u.min_t_mod);
			int max_t = upload_row_of(qword, u.max_t_mod);
			const int hi_cand = qword - stride * min_t;
			const int lo_cand = qword - stride * max_t;
			min_t = imax(min_t, (lo_cand * u.dxt) >> 16);
			max_t = imin(max_t, (hi_cand * u.dxt) >> 16);

			bool found = false;
			for (int t = max_t; t >= min_t; t--)
			{
				const int cand = qword - stride * t;
				const int cand_t = (cand * u.dxt) >> 16;
				if (cand + cand_t * stride == qword)
				{
					found = true;
					x_xor = (cand_t & 1) << 1;
					pixel = (cand << 2) + (pixel & 3);
				}
			}
			if (!found)
				return false;
		}
		ux = pixel;
		uy = 0;
	}
	else if (stride == 0)
	{
		// Every line lands on the same words; the last line wins.
		ux = pixel;
		uy = u.height - 1;
	}
	else
	{
		uy = upload_row_of(pixel, u.inv_tmem_stride_words);
		ux = pixel - uy * stride;
		if (uy >= u.height)
		{
			// Width exceeds stride: the tail of the last line spills forward.
			ux += stride * (uy - u.height + 1);
			uy = u.height - 1;
		}
	}

	int iter_offset = 0;
	if (u.tmem_size != u.vram_size)
	{
		if (u.vram_size - u.tmem_size == 1)
		{
			// RDRAM texels twice as wide as the tile's: writes mirror; pick the surviving half.
			iter_offset = (ux & ~3) * 4;
			if ((ux & ~3) + 2 < (u.vram_effective_width >> (3 - u.vram_size)))
				iter_offset += 8;
		}
		else if (u.tmem_size == 2 && u.vram_size == 1)
		{
			// 8bpp source into a 16bpp tile: only every other group of four words is written.
			if (ux & 4)
			{
				if ((stride & 4) != 0 && uy > 0)
				{
					uy--;
					ux += stride;
				}
				else
					return false;
			}
			iter_offset = ux & ~3;
		}
	}
	else
		iter_offset = (ux & ~3) * 2;

	if (ux >= u.width)
		return false;

	const int line_addr = u.vram_addr + ((uy * u.vram_width) << (u.vram_size - 1));
	ux ^= ((uy & 1) << 1) | x_xor;
	const uint32_t addr = uint32_t(line_addr + iter_offset + 2 * (ux & 3));
	value = rdram_read16(rdram, addr);
	return true;
}

// 32bpp and YUV tiles: the 64-bit load is split between lower and upper TMEM halves.
RDP_HD bool tmem_gather_32(const TmemUpload &u, const RdramView &rdram, int word_index, bool upper_half, bool yuv,
                           uint32_t &value)
{
	const int base = (u.tmem_offset & 0x7ff) >> 1;
	const int stride = u.tmem_stride_words;
	int pixel = (word_index - base) & 0x3ff;
	int ux, uy;
	int x_xor = 0;

	if (u.mode == UPLOAD_BLOCK)
	{
		const int qword = pixel >> 1;
		if (stride == 0)
		{
			const int first = qword & ~1;
			const int second = first + 1;
			const int first_t = (first * u.dxt) >> 16;
			const int second_t = (second * u.dxt) >> 16;
			if (first_t != second_t)
			{
				// The pair straddles a T increment; either, both or neither may land here.
				const int first_dst = first ^ (first_t & 1);
				const int second_dst = second ^ (second_t & 1);
				if (second_dst == qword)
					x_xor = (second_t & 1) << 1;
				else if (first_dst == qword)
					x_xor = (first_t & 1) << 1;
				else
					return false;
			}
			else
				x_xor = (first_t & 1) << 1;
		}
		else
		{
			int min_t = upload_row_of(qword & ~1, u.min_t_mod);
			int max_t = upload_row_of(qword | 1, u.max_t_mod);
			const int hi_cand = (qword | 1) - stride * min_t;
			const int lo_cand = (qword & ~1) - stride * max_t;
			min_t = imax(min_t, (lo_cand * u.dxt) >> 16);
			max_t = imin(max_t, (hi_cand * u.dxt) >> 16);

			bool found = false;
			for (int t = max_t; t >= min_t && !found; t--)
			{
				const int cand_a = (qword & ~1) - stride * t;
				const int cand_b = (qword | 1) - stride * t;
				const int t_a = (cand_a * u.dxt) >> 16;
				const int t_b = (cand_b * u.dxt) >> 16;
				if (((cand_b + t_b * stride) ^ (t_b & 1)) == qword)
				{
					found = true;
					pixel = (cand_b << 1) + (pixel & 1);
				}
				else if (((cand_a + t_a * stride) ^ (t_a & 1)) == qword)
				{
					found = true;
					pixel = (cand_a << 1) + (pixel & 1);
				}
			}
			if (!found)
				return false;
		}
		ux = pixel;
		uy = 0;
	}
	else if (stride == 0)
	{
		ux = pixel;
		uy = u.height - 1;
	}
	else
	{
		uy = upload_row_of(pixel, u.inv_tmem_stride_words);
		ux = pixel - uy * stride;
		if (uy >= u.height)
		{
			ux += stride * (uy - u.height + 1);
			uy = u.height - 1;
		}
	}

	const int last_line_x = ux ^ ((uy & 1) << 1);
	if (last_line_x >= u.width && uy > 0)
	{
		// Not reached by the last line; the line before may have written it.
		uy--;
		ux += stride;
	}

	ux ^= ((uy & 1) << 1) | x_xor;

	int iter_offset = 0;
	if (u.vram_size == 3 || yuv)
		iter_offset = 4 * (ux & ~1);
	else if (u.vram_size == 2)
	{
		if (ux & 2)
		{
			if (uy > 0)
			{
				uy--;
				ux += stride;
				ux ^= 2;
			}
			else
				return false;
		}
		iter_offset = 2 * (ux & ~1);
	}
	else if (u.vram_size == 1)
	{
		for (int i = 0; i < 4 && uy > 0 && (ux & 6) != 0; i++)
		{
			uy--;
			ux += stride;
			ux ^= 2;
		}
		if (ux & 6)
			return false;
		iter_offset = ux & ~1;
	}

	if (ux >= u.width)
		return false;

	const int line_addr = u.vram_addr + ((uy * u.vram_width) << (u.vram_size - 1));
	const uint32_t addr = uint32_t(line_addr + iter_offset + 4 * (ux & 1));
	uint32_t word = rdram_read32(rdram, addr);

	if (yuv)
	{
		if (upper_half)
			word = (((word >> 16) & 0xffu) << 8) | (word & 0xffu); // Y0 Y1
		else
			word = (((word >> 24) & 0xffu) << 8) | ((word >> 8) & 0xffu); // U V
	}
	else
		word = (word >> (upper_half ? 0u : 16u)) & 0xffffu;
	value = word;
	return true;
}

// LoadTLUT: each palette entry is replicated four times (one per TMEM bank).
RDP_HD bool tmem_gather_tlut(const TmemUpload &u, const RdramView &rdram, int word_index, uint32_t &value)
{
	const int base = (u.tmem_offset & 0xfff) >> 1;
	const int pixel = (word_index - base) & 0x7ff;
	const int size_delta = u.vram_size - u.tmem_size;
	int src;

	if (size_delta == 2)
	{
		src = (pixel >> 2) << (u.vram_size - 2);
	}
	else if (size_delta == 1)
	{
		if (pixel & 4)
			return false;
		src = (pixel & ~7) >> (u.tmem_size + (u.vram_size == 2 ? 2 : 0));
	}
	else if (size_delta == 0)
	{
		if (pixel & 0xc)
			return false;
		src = (pixel & ~3) >> (u.tmem_size + (u.vram_size == 2 ? 2 : 0));
	}
	else if (size_delta == -1)
	{
		if (pixel & 0x1c)
			return false;
		src = (pixel >> u.tmem_size) & ~7;
	}
	else
	{
		// 4bpp tile fed from 32bpp RDRAM: mirrored writes.
		int span_pixel = (pixel >> 2) * 4;
		if (span_pixel + 2 < u.vram_effective_width)
			span_pixel += 2;
		src = span_pixel;
	}
	if (src >= u.vram_effective_width)
		return false;

	uint32_t addr = uint32_t(u.vram_addr + (src << (u.vram_size - 1)));
	addr += 2u * (addr & 1u) * uint32_t(pixel & 3); // misaligned TLUT source quirk
	value = rdram_read16(rdram, addr);
	return true;
}

// Apply one upload to one TMEM word (`slot` is the storage index; the logical word is slot ^ 1).
RDP_HD void tmem_apply_upload(const TmemUpload &u, const RdramView &rdram, int slot, uint32_t &value)
{
	const int word_index = slot ^ 1;
	uint32_t v;
	bool hit;
	if (u.mode == UPLOAD_TLUT)
		hit = tmem_gather_tlut(u, rdram, word_index, v);
	else
	{
		const bool yuv = u.tmem_fmt == 1;
		if (u.tmem_size == 3 || yuv)
			hit = tmem_gather_32(u, rdram, word_index & 0x3ff, word_index >= 0x400, yuv, v);
		else
			hit = tmem_gather_16(u, rdram, word_index, v);
	}
	if (hit)
		value = v & 0xffffu;
}
}
