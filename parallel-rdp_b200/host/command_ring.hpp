// command_ring.hpp -- single-producer / single-consumer ring between the emulator thread and the command thread.
//
// Role mirrors parallel-rdp/command_ring.{hpp,cpp} (worker thread, 500 us idle => one idle kick) under the MIT
// licence of the reference (Copyright (c) 2020 Themaister); the class name and the public method names are the
// reference's so that RDP::CommandProcessor reads the same.
// The reference takes a mutex and signals a condition variable on every command; at the 1e8 commands/s this
// project targets that alone exceeds the per-command budget, so this ring is lock-free and record-oriented:
//   * a record is [header, payload words...] and never wraps -- when the tail of the buffer is too short the
//     producer writes a PAD header and restarts at offset 0 -- so the consumer executes commands *in place*,
//     without copying them out of the ring;
//   * a record is either one command (enqueue_command), a run of back-to-back commands (enqueue_stream), a
//     {pointer, length} pair naming a caller-owned run (enqueue_stream_borrowed), or a *meta* record
//     (enqueue_meta: timeline signal, flush, quirks, queued scan-out).  Meta records are a record kind of their
//     own, not in-band opcodes: nothing a guest display list or a dump file contains can turn into one;
//   * publishing is one release store; the mutex / condvar pair is only touched when one side has to sleep.
//
// Waiting never burns a core: either side spins for a few microseconds (RDP_B200_RING_SPIN_US, default 20) and
// then blocks on its condition variable.  An idle consumer wakes once, 500 us after the last record, to tell the
// processor the producer went quiet (the reference's MetaIdle, command_ring.cpp:89-134) and then sleeps until
// the next publish -- a farm of idle processors costs no CPU time.
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <stddef.h>
#include <stdint.h>
#include <thread>
#include <vector>

namespace RDP
{
class CommandProcessor;

// Meta records (the reference's Op::Meta* commands, rdp_common.hpp:37-45, plus the queued scan-out).
enum class RingMeta : uint32_t
{
	SignalTimeline = 1, // payload: value lo, value hi
	Flush = 2,
	Idle = 3,
	SetQuirks = 4,      // payload: quirk word
	Scanout = 5,        // payload: pointer lo, pointer hi of a heap-allocated request the command thread consumes
};

class CommandRing
{
public:
	~CommandRing();
	void init(CommandProcessor *processor, unsigned count_words);
	bool is_running() const { return thr.joinable(); }
	void enqueue_command(unsigned num_words, const uint32_t *words);
	// `words` holds whole commands back to back (the caller guarantees it ends on a command boundary).
	void enqueue_stream(size_t num_words, const uint32_t *words);
	// Zero-copy variant: only {pointer, length} cross the ring; the consumer reads the caller's buffer directly.
	void enqueue_stream_borrowed(size_t num_words, const uint32_t *words);
	void enqueue_meta(RingMeta kind, unsigned num_words, const uint32_t *payload);
	void drain();
	// Blocks the producer until `flag` (written by the command thread while it executes a record) reaches `value`.
	void wait_until_at_least(const std::atomic<uint64_t> &flag, uint64_t value);
	void shutdown() { teardown_thread(); }
	// Pins the command thread to `cpus` (ignored if empty or unsupported).
	void set_affinity(const std::vector<int> &cpus);

private:
	enum : uint32_t
	{
		kHeaderStream = 1u << 31, // payload is a run of commands
		kHeaderBorrowed = 1u << 30, // payload is {pointer lo, pointer hi, words lo, words hi} of a caller-owned run
		kHeaderMeta = 1u << 29, // payload[0] = RingMeta, then its arguments
		kHeaderFlags = kHeaderStream | kHeaderBorrowed | kHeaderMeta,
		kHeaderPad = 0xffffffffu, // skip to the start of the buffer
		kHeaderExit = 0x1fffffffu,
		kMaxRecordWords = 32 * 1024, // stream records are cut to this so the consumer starts early
	};

	CommandProcessor *processor = nullptr;
	std::thread thr;
	std::vector<uint32_t> ring;
	uint64_t size = 0;
	int64_t spin_ns = 20000;

	alignas(64) std::atomic<uint64_t> write_count{0};     // published by the producer (monotonic word count)
	alignas(64) std::atomic<uint64_t> read_count{0};      // space released by the consumer = commands executed
	alignas(64) std::atomic<bool> consumer_asleep{false};
	std::atomic<bool> producer_asleep{false};
	uint64_t producer_pos = 0; // producer-private copy of write_count

	std::mutex lock;
	std::condition_variable cond_consumer;
	std::condition_variable cond_producer;

	uint32_t *reserve(uint32_t payload_words, uint32_t header);
	void publish(uint32_t payload_words);
	void thread_loop();
	void teardown_thread();
	template <class Pred>
	void producer_wait(Pred pred);
};
}
