// rdp_luts.cpp -- host-side generators for the two look-up tables the pixel paths need.
//
// The reference ships them as literal data (parallel-rdp/luts.hpp: blender_lut[0x8000], gamma_table[256+256*64]),
// which in turn were produced by the oracle's init code (angrylion rdp/blender.c:304 blender_init_lut,
// vi/gamma.c:63 vi_gamma_init).  We regenerate them from the underlying rule and pin the bytes with a hash
// test (tests/test_luts.py) instead of carrying 1,500 lines of numbers.
#include <stdint.h>

namespace rdpb
{
// The RDP blender normalises (p*a + m*b) / (a + b) with a small bit-serial non-restoring divider:
// a 3-bit partial remainder, the 4-bit divisor d (= sum of the two 3-bit weights + 1) and eight quotient steps.
// Index = (d << 11) | n where n is the 11-bit numerator.
void build_blender_lut(uint8_t *lut)
{
	for (int index = 0; index < 0x8000; index++)
	{
		const int d = (index >> 11) & 0xf;
		const int n = index & 0x7ff;
		const int not_d = ~d & 0xf;

		int quotient = 0;
		int partial = (not_d + (n >> 8) + 1) & 7;
		bool previous_bit = false; // quotient bit produced by the previous step
		for (int step = 0; step < 8; step++)
		{
			const int next_bit = (n >> (7 - step)) & 1;
			// Subtract (add two's complement) after a 1, add back after a 0.
			const int t = previous_bit ? (not_d + (partial << 1) + next_bit + 1) : (d + (partial << 1) + next_bit);
			partial = t & 7;
			previous_bit = (t & 0x10) != 0;
			if (previous_bit)
				quotient |= 1 << (7 - step);
		}
		lut[index] = uint8_t(quotient);
	}
}

// The VI's gamma ROM holds floor(sqrt(x)) of a 14-bit input (7 result bits).
static int vi_isqrt(int a)
{
	int r = 0;
	while ((r + 1) * (r + 1) <= a)
		r++;
	return r;
}

// gamma[0..255]           : plain gamma,   2*isqrt(c << 6)
// gamma[256 + (c<<6 | d)] : dithered gamma, 2*isqrt(c<<6 | d) for a 6-bit dither value d
void build_gamma_lut(uint8_t *lut)
{
	for (int i = 0; i < 256; i++)
		lut[i] = uint8_t(vi_isqrt(i << 6) << 1);
	for (int i = 0; i < 0x4000; i++)
		lut[256 + i] = uint8_t(vi_isqrt(i) << 1);
}
}
