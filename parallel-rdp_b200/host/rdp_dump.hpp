// rdp_dump.hpp -- RDPDUMP2 command-stream dumps: reader (DumpPlayer) and writer (RDPDumpWriter).
//
// The file format is the reference's (ref: rdp_dump.cpp:29-185 reader, parallel-rdp/rdp_dump_write.cpp:34-151 writer):
//   "RDPDUMP2", u32 rdram_size, u32 hidden_rdram_size, then records, each starting with a u32 command id:
//     1 UPDATE_DRAM          u32 offset, u32 size, size bytes     patch of the RDRAM image
//     7 UPDATE_DRAM_FLUSH                                         the patched image is now current -> listener
//     8 UPDATE_HIDDEN_DRAM   u32 offset, u32 size, size bytes
//     9 UPDATE_HIDDEN_DRAM_FLUSH
//     2 RDP_COMMAND          u32 op, u32 num_words, words
//     3 SET_VI_REGISTER      u32 index, u32 value
//     4 END_FRAME     5 SIGNAL_COMPLETE     6 EOF
// Everything is little-endian host order, as the reference writes it with fwrite.
//
// This implementation differs from the reference's in how it gets there: the reader parses the whole file from
// one buffer and hands out events (iterate() delivers them to a listener with the reference's interface), and the
// writer merges runs of adjacent dirty 4 KiB blocks into one UPDATE record (the reader side, ours or the
// reference's, accepts any offset/size inside the image).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

namespace RDP
{
enum class DumpCommand : uint32_t
{
	Invalid = 0,
	UpdateDram = 1,
	RDPCommand = 2,
	SetVIRegister = 3,
	EndFrame = 4,
	SignalComplete = 5,
	EndOfFile = 6,
	UpdateDramFlush = 7,
	UpdateHiddenDram = 8,
	UpdateHiddenDramFlush = 9,
};

// ref: rdp_dump.hpp:34-44 (Op / VIRegister are plain integers at this boundary).
struct CommandListenerInterface
{
	virtual ~CommandListenerInterface() = default;
	virtual void set_vi_register(uint32_t reg, uint32_t value) = 0;
	virtual void signal_complete() = 0;
	virtual void command(uint32_t op, uint32_t num_words, const uint32_t *words) = 0;
	virtual void end_frame() = 0;
	virtual void eof() = 0;
	virtual void update_rdram(const void *data, size_t size, size_t offset) = 0;
	virtual void update_hidden_rdram(const void *data, size_t size, size_t offset) = 0;
};

// One decoded record, as next_event() returns it.  UPDATE_DRAM / UPDATE_HIDDEN_DRAM records are applied to the
// player's images and never surface; the FLUSH records carry the whole current image.
struct DumpEvent
{
	DumpCommand kind = DumpCommand::Invalid;
	uint32_t arg0 = 0, arg1 = 0;   // RDPCommand: op, -;  SetVIRegister: index, value
	const uint32_t *words = nullptr; // RDPCommand
	uint32_t num_words = 0;
	const uint8_t *data = nullptr;   // *Flush: the image
	size_t size = 0;
};

class DumpPlayer
{
public:
	bool load_dump(const char *path);
	bool load_memory(const void *data, size_t size); // takes a copy
	size_t get_rdram_size() const { return rdram.size(); }
	size_t get_hidden_rdram_size() const { return hidden.size(); }
	void set_command_interface(CommandListenerInterface *iface_) { iface = iface_; }
	// Decode the next record.  False at end of file, on a truncated or malformed record.
	bool next_event(DumpEvent &ev);
	// Reference-style stepping: deliver the next record to the listener; false once the stream ends (after eof()).
	bool iterate();
	bool rewind();

private:
	CommandListenerInterface *iface = nullptr;
	std::vector<uint8_t> file;
	size_t cursor = 0;
	std::vector<uint8_t> rdram, hidden;
	std::vector<uint32_t> words;
	bool read_u32(uint32_t &v);
	bool apply_patch(std::vector<uint8_t> &image);
};

class RDPDumpWriter
{
public:
	~RDPDumpWriter();
	bool init(const char *path, uint32_t dram_size, uint32_t hidden_dram_size);
	void flush_dram(const void *dram, uint32_t size);
	void flush_hidden_dram(const void *dram, uint32_t size);
	void signal_complete();
	void emit_command(uint32_t command, const uint32_t *cmd_data, uint32_t cmd_words);
	void set_vi_register(uint32_t vi_register, uint32_t value);
	void end_frame();
	void end(); // writes EOF and closes; the destructor does it too

private:
	FILE *fp = nullptr;
	std::vector<uint8_t> dram_shadow, hidden_shadow;
	void put_u32(uint32_t v);
	void flush_image(const uint8_t *now, uint32_t size, std::vector<uint8_t> &shadow, DumpCommand patch, DumpCommand done);
};
}
