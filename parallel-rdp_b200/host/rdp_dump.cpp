// rdp_dump.cpp -- see rdp_dump.hpp.
#include "rdp_dump.hpp"
#include <algorithm>
#include <string.h>

namespace RDP
{
static const char kMagic[8] = { 'R', 'D', 'P', 'D', 'U', 'M', 'P', '2' };
static const size_t kHeaderBytes = 16;

// ---- reader --------------------------------------------------------------------------------------------------
bool DumpPlayer::load_memory(const void *data, size_t size)
{
	file.assign(static_cast<const uint8_t *>(data), static_cast<const uint8_t *>(data) + size);
	if (file.size() < kHeaderBytes || memcmp(file.data(), kMagic, sizeof(kMagic)) != 0)
		return false;
	uint32_t rdram_size, hidden_size;
	memcpy(&rdram_size, file.data() + 8, 4);
	memcpy(&hidden_size, file.data() + 12, 4);
	// Same acceptance rule as the reference reader (ref: rdp_dump.cpp:60-63).
	if (rdram_size != 4u * 1024 * 1024 && rdram_size != 8u * 1024 * 1024)
		return false;
	if (hidden_size != 4u * 1024 * 1024)
		return false;
	rdram.assign(rdram_size, 0);
	hidden.assign(hidden_size, 0);
	cursor = kHeaderBytes;
	return true;
}

bool DumpPlayer::load_dump(const char *path)
{
	FILE *f = fopen(path, "rb");
	if (!f)
		return false;
	std::vector<uint8_t> bytes;
	uint8_t chunk[1 << 16];
	size_t n;
	while ((n = fread(chunk, 1, sizeof(chunk), f)) != 0)
		bytes.insert(bytes.end(), chunk, chunk + n);
	fclose(f);
	return load_memory(bytes.data(), bytes.size());
}

bool DumpPlayer::rewind()
{
	if (file.size() < kHeaderBytes)
		return false;
	cursor = kHeaderBytes;
	std::fill(rdram.begin(), rdram.end(), 0);
	std::fill(hidden.begin(), hidden.end(), 0);
	return true;
}

bool DumpPlayer::read_u32(uint32_t &v)
{
	if (file.size() - cursor < 4)
		return false;
	memcpy(&v, file.data() + cursor, 4);
	cursor += 4;
	return true;
}

bool DumpPlayer::apply_patch(std::vector<uint8_t> &image)
{
	uint32_t offset, size;
	if (!read_u32(offset) || !read_u32(size))
		return false;
	if (uint64_t(offset) + size > image.size() || file.size() - cursor < size)
		return false;
	memcpy(image.data() + offset, file.data() + cursor, size);
	cursor += size;
	return true;
}

bool DumpPlayer::next_event(DumpEvent &ev)
{
	for (;;)
	{
		uint32_t id;
		if (!read_u32(id))
			return false;
		ev = DumpEvent();
		ev.kind = DumpCommand(id);
		switch (ev.kind)
		{
		case DumpCommand::UpdateDram:
			if (!apply_patch(rdram))
				return false;
			continue;
		case DumpCommand::UpdateHiddenDram:
			if (!apply_patch(hidden))
				return false;
			continue;
		case DumpCommand::UpdateDramFlush:
			ev.data = rdram.data();
			ev.size = rdram.size();
			return true;
		case DumpCommand::UpdateHiddenDramFlush:
			ev.data = hidden.data();
			ev.size = hidden.size();
			return true;
		case DumpCommand::RDPCommand:
		{
			uint32_t count;
			if (!read_u32(ev.arg0) || !read_u32(count) || (file.size() - cursor) / 4 < count)
				return false;
			words.resize(count);
			if (count)
				memcpy(words.data(), file.data() + cursor, size_t(count) * 4);
			cursor += size_t(count) * 4;
			ev.words = words.data();
			ev.num_words = count;
			return true;
		}
		case DumpCommand::SetVIRegister:
			return read_u32(ev.arg0) && read_u32(ev.arg1);
		case DumpCommand::EndFrame:
		case DumpCommand::SignalComplete:
		case DumpCommand::EndOfFile:
			return true;
		default:
			return false;
		}
	}
}

bool DumpPlayer::iterate()
{
	DumpEvent ev;
	if (!next_event(ev))
		return false;
	switch (ev.kind)
	{
	case DumpCommand::EndOfFile:
		iface->eof();
		return false;
	case DumpCommand::SetVIRegister:
		iface->set_vi_register(ev.arg0, ev.arg1);
		break;
	case DumpCommand::RDPCommand:
		iface->command(ev.arg0, ev.num_words, ev.words);
		break;
	case DumpCommand::EndFrame:
		iface->end_frame();
		break;
	case DumpCommand::SignalComplete:
		iface->signal_complete();
		break;
	case DumpCommand::UpdateDramFlush:
		iface->update_rdram(ev.data, ev.size, 0);
		break;
	case DumpCommand::UpdateHiddenDramFlush:
		iface->update_hidden_rdram(ev.data, ev.size, 0);
		break;
	default:
		return false;
	}
	return true;
}

// ---- writer --------------------------------------------------------------------------------------------------
RDPDumpWriter::~RDPDumpWriter() { end(); }

void RDPDumpWriter::put_u32(uint32_t v) { fwrite(&v, 4, 1, fp); }

bool RDPDumpWriter::init(const char *path, uint32_t dram_size, uint32_t hidden_dram_size)
{
	if (fp)
		return false;
	fp = fopen(path, "wb");
	if (!fp)
		return false;
	dram_shadow.assign(dram_size, 0);
	hidden_shadow.assign(hidden_dram_size, 0);
	fwrite(kMagic, 1, sizeof(kMagic), fp);
	put_u32(dram_size);
	put_u32(hidden_dram_size);
	return true;
}

void RDPDumpWriter::end()
{
	if (!fp)
		return;
	put_u32(uint32_t(DumpCommand::EndOfFile));
	fclose(fp);
	fp = nullptr;
	dram_shadow.clear();
	hidden_shadow.clear();
}

// Diff against the shadow copy at 4 KiB granularity (the reference's block size), one record per run of dirty blocks.
void RDPDumpWriter::flush_image(const uint8_t *now, uint32_t size, std::vector<uint8_t> &shadow, DumpCommand patch, DumpCommand done)
{
	if (!fp)
		return;
	const uint32_t block = 4096;
	size = std::min<uint32_t>(size, uint32_t(shadow.size()));
	uint32_t run_start = 0;
	bool in_run = false;
	for (uint32_t at = 0;; at += block)
	{
		const uint32_t len = at < size ? std::min(block, size - at) : 0;
		const bool dirty = len != 0 && memcmp(now + at, shadow.data() + at, len) != 0;
		if (dirty && !in_run)
		{
			run_start = at;
			in_run = true;
		}
		else if (!dirty && in_run)
		{
			const uint32_t end = std::min(at, size);
			put_u32(uint32_t(patch));
			put_u32(run_start);
			put_u32(end - run_start);
			fwrite(now + run_start, 1, end - run_start, fp);
			memcpy(shadow.data() + run_start, now + run_start, end - run_start);
			in_run = false;
		}
		if (len == 0)
			break;
	}
	put_u32(uint32_t(done));
}

void RDPDumpWriter::flush_dram(const void *dram, uint32_t size)
{
	flush_image(static_cast<const uint8_t *>(dram), size, dram_shadow, DumpCommand::UpdateDram, DumpCommand::UpdateDramFlush);
}

void RDPDumpWriter::flush_hidden_dram(const void *dram, uint32_t size)
{
	flush_image(static_cast<const uint8_t *>(dram), size, hidden_shadow, DumpCommand::UpdateHiddenDram, DumpCommand::UpdateHiddenDramFlush);
}

void RDPDumpWriter::signal_complete()
{
	if (fp)
	{
		put_u32(uint32_t(DumpCommand::SignalComplete));
		fflush(fp); // a dump is a debugging aid: keep it usable when the process dies before the writer is closed
	}
}

void RDPDumpWriter::emit_command(uint32_t command, const uint32_t *cmd_data, uint32_t cmd_words)
{
	if (!fp)
		return;
	put_u32(uint32_t(DumpCommand::RDPCommand));
	put_u32(command);
	put_u32(cmd_words);
	fwrite(cmd_data, 4, cmd_words, fp);
}

void RDPDumpWriter::set_vi_register(uint32_t vi_register, uint32_t value)
{
	if (!fp)
		return;
	put_u32(uint32_t(DumpCommand::SetVIRegister));
	put_u32(vi_register);
	put_u32(value);
}

void RDPDumpWriter::end_frame()
{
	if (fp)
	{
		put_u32(uint32_t(DumpCommand::EndFrame));
		fflush(fp);
	}
}
}
