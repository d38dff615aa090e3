// replayer_driver_b200.cpp -- the reference harness's "GPU" replayer on top of librdp_b200.so.
//
// This is the binding INTEGRATION.md describes, compiled for real: it implements the reference's own
// RDP::ReplayerDriver interface (replayer_driver.hpp:83-100) and its factory create_replayer_driver_parallel()
// (replayer_driver.hpp:143-145), method for method what replayer_driver_parallel.cpp:30-206 does with
// RDP::CommandProcessor, through the C-ABI of include/rdp_b200.h.  Linked with the reference's unmodified
// rdp_conformance.cpp / vi_conformance.cpp / rdp_command_builder.cpp / replayer_driver_angrylion.cpp it yields the
// reference's conformance programs -- same std::mt19937{1337} streams, same comparisons -- with this library in
// the GPU seat.  The library is loaded at run time (RDP_B200_LIBRARY, default: the product next to this tree), so
// the same binary also drives the test-only host simulator in the CPU suite.  Test infrastructure (oracle/).
#include "replayer_driver.hpp"
#include "rdp_b200.h"
#include "logging.hpp"

#include <dlfcn.h>
#include <stdexcept>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace RDP
{
namespace
{
struct Api
{
#define RDP_B200_FN(name) decltype(&rdp_b200_##name) name = nullptr
	RDP_B200_FN(create); RDP_B200_FN(destroy); RDP_B200_FN(device_is_supported); RDP_B200_FN(enqueue_command);
	RDP_B200_FN(flush); RDP_B200_FN(idle); RDP_B200_FN(begin_read_rdram); RDP_B200_FN(end_write_rdram);
	RDP_B200_FN(begin_read_hidden_rdram); RDP_B200_FN(end_write_hidden_rdram); RDP_B200_FN(get_rdram_size);
	RDP_B200_FN(get_hidden_rdram_size); RDP_B200_FN(get_tmem); RDP_B200_FN(set_vi_register);
	RDP_B200_FN(begin_vi_register_per_scanline); RDP_B200_FN(set_vi_register_for_scanline);
	RDP_B200_FN(latch_vi_register_for_scanline); RDP_B200_FN(end_vi_register_per_scanline); RDP_B200_FN(scanout_sync);
	RDP_B200_FN(default_scanout_options); RDP_B200_FN(set_validation_callback);
#undef RDP_B200_FN
	bool load()
	{
		std::string path;
		if (const char *env = getenv("RDP_B200_LIBRARY"))
			path = env;
		else
		{
			// oracle/_ref/<binary> -> ../../parallel-rdp_b200/librdp_b200.so
			Dl_info info;
			if (dladdr(reinterpret_cast<void *>(&create_replayer_driver_parallel), &info) && info.dli_fname)
			{
				path = info.dli_fname;
				const size_t slash = path.rfind('/');
				path = (slash == std::string::npos ? std::string(".") : path.substr(0, slash)) + "/../../parallel-rdp_b200/librdp_b200.so";
			}
		}
		void *h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
		if (!h)
		{
			LOGE("Cannot load %s: %s\n", path.c_str(), dlerror());
			return false;
		}
		bool ok = true;
#define BIND(name) ok = ok && (name = reinterpret_cast<decltype(name)>(dlsym(h, "rdp_b200_" #name))) != nullptr
		BIND(create); BIND(destroy); BIND(device_is_supported); BIND(enqueue_command); BIND(flush); BIND(idle);
		BIND(begin_read_rdram); BIND(end_write_rdram); BIND(begin_read_hidden_rdram); BIND(end_write_hidden_rdram);
		BIND(get_rdram_size); BIND(get_hidden_rdram_size); BIND(get_tmem); BIND(set_vi_register);
		BIND(begin_vi_register_per_scanline); BIND(set_vi_register_for_scanline); BIND(latch_vi_register_for_scanline);
		BIND(end_vi_register_per_scanline); BIND(scanout_sync); BIND(default_scanout_options); BIND(set_validation_callback);
#undef BIND
		return ok;
	}
};

Api &api()
{
	static Api a;
	static bool loaded = a.load();
	if (!loaded)
		throw std::runtime_error("librdp_b200.so is not available.");
	return a;
}

class B200Replayer : public ReplayerDriver
{
public:
	B200Replayer(CommandInterface &player_, ReplayerEventInterface &iface_, bool upscale)
		: player(player_), iface(iface_)
	{
		// 64 KiB alignment as the reference asks of imported host memory (replayer_driver_parallel.cpp:37).
		if (posix_memalign(&host_memory, 64 * 1024, player.get_rdram_size()) != 0)
			throw std::bad_alloc();
		memset(host_memory, 0, player.get_rdram_size());
		const uint32_t flags = RDP_B200_FLAG_HOST_VISIBLE_HIDDEN_RDRAM | RDP_B200_FLAG_HOST_VISIBLE_TMEM |
		                       (upscale ? RDP_B200_FLAG_UPSCALING_2X : 0);
		const char *dev = getenv("RDP_B200_CUDA_DEVICE");
		// PARALLEL_RDP_DUMP_PATH names one file: keep it for the native-resolution replayer (the harness also builds a 2x one).
		std::string dump_path;
		if (upscale && getenv("PARALLEL_RDP_DUMP_PATH"))
		{
			dump_path = getenv("PARALLEL_RDP_DUMP_PATH");
			unsetenv("PARALLEL_RDP_DUMP_PATH");
		}
		gpu = api().create(dev ? atoi(dev) : 0, host_memory, 0, player.get_rdram_size(), player.get_hidden_rdram_size(), flags);
		if (!dump_path.empty())
			setenv("PARALLEL_RDP_DUMP_PATH", dump_path.c_str(), 1);
		if (!gpu || !api().device_is_supported(gpu))
			throw std::runtime_error("GPU is not supported.");
		api().set_validation_callback(gpu, [](void *, int, const char *msg) {
			LOGE("RDP crash: %s\n", msg);
			abort();
		}, nullptr);
	}

	~B200Replayer() override
	{
		if (gpu)
			api().destroy(gpu);
		free(host_memory);
	}

private:
	CommandInterface &player;
	ReplayerEventInterface &iface;
	void *host_memory = nullptr;
	rdp_b200_processor *gpu = nullptr;
	struct { unsigned left = 0, right = 0, top = 0, bottom = 0; bool enable = false; } crop_rect;

	void eof() override { iface.eof(); }
	void signal_complete() override
	{
		api().flush(gpu);
		iface.signal_complete();
	}
	void update_rdram(const void *data, size_t size, size_t offset) override
	{
		api().idle(gpu);
		memcpy(static_cast<uint8_t *>(host_memory) + offset, data, size);
		api().end_write_rdram(gpu);
	}
	void update_hidden_rdram(const void *data, size_t size, size_t offset) override
	{
		api().idle(gpu);
		memcpy(static_cast<uint8_t *>(api().begin_read_hidden_rdram(gpu)) + offset, data, size);
		api().end_write_hidden_rdram(gpu);
	}
	void command(Op command_id, uint32_t num_words, const uint32_t *words) override
	{
		api().enqueue_command(gpu, num_words, words);
		iface.notify_command(command_id, num_words, words);
	}
	void end_frame() override
	{
		rdp_b200_scanout_options opts;
		api().default_scanout_options(&opts);
		opts.blend_previous_frame = 1;
		opts.upscale_deinterlacing = 0;
		opts.crop_left = crop_rect.left;
		opts.crop_right = crop_rect.right;
		opts.crop_top = crop_rect.top;
		opts.crop_bottom = crop_rect.bottom;
		opts.crop_enable = crop_rect.enable ? 1 : 0;
		std::vector<uint32_t> colors(size_t(640) * 576 * 64);
		unsigned width = 0, height = 0;
		api().scanout_sync(gpu, &opts, colors.data(), colors.size(), &width, &height);
		iface.update_screen(colors.data(), width, height, width);
		crop_rect.enable = false;
	}
	void set_vi_register(VIRegister index, uint32_t value) override { api().set_vi_register(gpu, uint32_t(index), value); }
	uint8_t *get_rdram() override
	{
		api().idle(gpu);
		return static_cast<uint8_t *>(host_memory);
	}
	size_t get_rdram_size() override { return api().get_rdram_size(gpu); }
	uint8_t *get_hidden_rdram() override
	{
		api().idle(gpu);
		return static_cast<uint8_t *>(api().begin_read_hidden_rdram(gpu));
	}
	size_t get_hidden_rdram_size() override { return api().get_hidden_rdram_size(gpu); }
	uint8_t *get_tmem() override
	{
		api().idle(gpu);
		return static_cast<uint8_t *>(api().get_tmem(gpu));
	}
	void idle() override { api().idle(gpu); }
	void flush_caches() override
	{
		api().end_write_rdram(gpu);
		api().end_write_hidden_rdram(gpu);
	}
	void invalidate_caches() override
	{
		api().begin_read_rdram(gpu);
		api().begin_read_hidden_rdram(gpu);
	}
	void begin_vi_register_per_scanline() override { api().begin_vi_register_per_scanline(gpu, 3u); } // HSTART | XSCALE
	void set_vi_register_for_scanline(unsigned vi_line, uint32_t h_start, uint32_t x_scale) override
	{
		api().set_vi_register_for_scanline(gpu, 1u, h_start);
		api().set_vi_register_for_scanline(gpu, 2u, x_scale);
		api().latch_vi_register_for_scanline(gpu, vi_line);
	}
	void end_vi_register_per_scanline() override { api().end_vi_register_per_scanline(gpu); }
	void set_crop_rect(unsigned left, unsigned right, unsigned top, unsigned bottom) override
	{
		crop_rect.left = left;
		crop_rect.right = right;
		crop_rect.top = top;
		crop_rect.bottom = bottom;
		crop_rect.enable = true;
	}
};
}

std::unique_ptr<ReplayerDriver> create_replayer_driver_parallel(Vulkan::Device &, CommandInterface &player, ReplayerEventInterface &iface,
                                                                bool, bool upscale)
{
	return std::make_unique<B200Replayer>(player, iface, upscale);
}
}
