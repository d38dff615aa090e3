// ref_dump_shim.cpp -- flat C entry points over the reference's own RDPDUMP2 reader and writer, built from the
// sources where they lie under /root/reference (rdp_dump.cpp, parallel-rdp/rdp_dump_write.cpp) into
// oracle/_ref/libref_dump.so.  TEST INFRASTRUCTURE ONLY: the product's reader / writer
// (parallel-rdp_b200/host/rdp_dump.cpp) are checked against these for file-format interchangeability.
#include "rdp_dump.hpp"       // reference, /root/reference
#include "rdp_dump_write.hpp" // reference, /root/reference/parallel-rdp
#include <stddef.h>
#include <stdint.h>

namespace
{
typedef void (*refdump_cb)(void *user, uint32_t kind, uint32_t a, uint32_t b, const void *data, size_t size);

struct Listener : RDP::CommandListenerInterface
{
	refdump_cb cb;
	void *user;
	long events = 0;
	void emit(uint32_t kind, uint32_t a, uint32_t b, const void *data, size_t size)
	{
		events++;
		cb(user, kind, a, b, data, size);
	}
	void set_vi_register(RDP::VIRegister reg, uint32_t value) override { emit(3, uint32_t(reg), value, nullptr, 0); }
	void signal_complete() override { emit(5, 0, 0, nullptr, 0); }
	void command(RDP::Op id, uint32_t n, const uint32_t *words) override { emit(2, uint32_t(id), n, words, size_t(n) * 4); }
	void end_frame() override { emit(4, 0, 0, nullptr, 0); }
	void eof() override { emit(6, 0, 0, nullptr, 0); }
	void update_rdram(const void *data, size_t size, size_t) override { emit(7, 0, 0, data, size); }
	void update_hidden_rdram(const void *data, size_t size, size_t) override { emit(9, 0, 0, data, size); }
};
}

extern "C" {
// Plays the whole file through the reference DumpPlayer; returns the number of events delivered, -1 if it does not load.
long refdump_play(const char *path, refdump_cb cb, void *user)
{
	RDP::DumpPlayer player;
	if (!player.load_dump(path))
		return -1;
	Listener l;
	l.cb = cb;
	l.user = user;
	player.set_command_interface(&l);
	while (player.iterate())
	{
	}
	return l.events;
}

void *refdump_writer_open(const char *path, uint32_t rdram_size, uint32_t hidden_size)
{
	auto *w = new RDP::RDPDumpWriter;
	if (!w->init(path, rdram_size, hidden_size))
	{
		delete w;
		return nullptr;
	}
	return w;
}
void refdump_writer_flush_dram(void *w, const void *p, uint32_t n) { static_cast<RDP::RDPDumpWriter *>(w)->flush_dram(p, n); }
void refdump_writer_flush_hidden_dram(void *w, const void *p, uint32_t n) { static_cast<RDP::RDPDumpWriter *>(w)->flush_hidden_dram(p, n); }
void refdump_writer_command(void *w, uint32_t op, const uint32_t *words, uint32_t n) { static_cast<RDP::RDPDumpWriter *>(w)->emit_command(op, words, n); }
void refdump_writer_set_vi_register(void *w, uint32_t reg, uint32_t v) { static_cast<RDP::RDPDumpWriter *>(w)->set_vi_register(reg, v); }
void refdump_writer_end_frame(void *w) { static_cast<RDP::RDPDumpWriter *>(w)->end_frame(); }
void refdump_writer_signal_complete(void *w) { static_cast<RDP::RDPDumpWriter *>(w)->signal_complete(); }
void refdump_writer_close(void *w) { delete static_cast<RDP::RDPDumpWriter *>(w); }
}
