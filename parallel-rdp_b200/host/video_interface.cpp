// video_interface.cpp -- see video_interface.hpp.
#include "video_interface.hpp"
#include <string.h>

namespace RDP
{
void VideoInterface::init(rdpb_device *device_, size_t rdram_size_)
{
	device = device_;
	rdram_size = rdram_size_;
}

void VideoInterface::set_vi_register(VIRegister reg, uint32_t value)
{
	vi_registers[unsigned(reg)] = value;
}

void VideoInterface::begin_vi_register_per_scanline(PerScanlineRegisterFlags) {}
void VideoInterface::set_vi_register_for_scanline(PerScanlineRegisterBits, uint32_t) {}
void VideoInterface::latch_vi_register_for_scanline(unsigned) {}
void VideoInterface::end_vi_register_per_scanline() {}

ScanoutImage VideoInterface::scanout(const ScanoutOptions &)
{
	return {};
}
}
