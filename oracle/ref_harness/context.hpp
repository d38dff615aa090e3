// Shim for Granite's vulkan/context.hpp: the reference's conformance harness (rdp_conformance.cpp, vi_conformance.cpp,
// conformance_utils.hpp) is compiled UNMODIFIED from /root/reference against these few stand-ins, because the only
// thing it uses a Vulkan context / device for is to construct its GPU replayer -- which here is
// replayer_driver_b200.cpp over librdp_b200.so.  Test infrastructure (oracle/); nothing in the product includes it.
#pragma once
#include <stdint.h>

namespace Granite
{
class Filesystem;
class ThreadGroup;
class TimelineTraceFile;
}

namespace Vulkan
{
class Context
{
public:
	struct SystemHandles
	{
		Granite::Filesystem *filesystem = nullptr;
		Granite::ThreadGroup *thread_group = nullptr;
		Granite::TimelineTraceFile *timeline_trace_file = nullptr;
	};
	static bool init_loader(void *) { return true; }
	void set_system_handles(const SystemHandles &) {}
	bool init_instance_and_device(const char **, uint32_t, const char **, uint32_t) { return true; }
};
}
