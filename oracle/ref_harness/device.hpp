// Shim for Granite's vulkan/device.hpp (see context.hpp).
#pragma once
#include "context.hpp"

namespace Vulkan
{
class Device
{
public:
	void set_context(const Context &) {}
	void wait_idle() {}
	void next_frame_context() {}
	void begin_renderdoc_capture() {}
	void end_renderdoc_capture() {}
	static bool init_renderdoc_capture() { return false; }
};
}
