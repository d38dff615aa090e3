// command_ring.hpp -- single-producer / single-consumer ring between the emulator thread and the command thread.
//
// Role mirrors parallel-rdp/command_ring.{hpp,cpp} (4096-word ring, worker thread, 500 us idle => MetaIdle).
// The reference takes a mutex and signals a condition variable on every command; at the 1e8 commands/s this
// project targets that alone exceeds the per-command budget, so this ring is lock-free and record-oriented:
//   * a record is [header, payload words...] and never wraps -- when the tail of the buffer is too short the
//     producer writes a PAD header and restarts at offset 0 -- so the consumer executes commands *in place*,
//     without copying them out of the ring;
//   * a record is either one command (enqueue_command) or a run of back-to-back commands (enqueue_stream),
//     which the consumer walks by opcode length: a whole frame costs a handful of publishes, not one per command;
//   * publishing is one release store; the mutex/condvar pair is only touched when one side has to sleep.
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
	void drain();
	void shutdown() { teardown_thread(); }

private:
	enum : uint32_t
	{
		kHeaderStream = 1u << 31, // payload is a run of commands
		kHeaderBorrowed = 1u << 30, // payload is {pointer lo, pointer hi, words lo, words hi} of a caller-owned run
		kHeaderPad = 0xffffffffu, // skip to the start of the buffer
		kHeaderExit = 0x7fffffffu,
		kMaxRecordWords = 32 * 1024, // stream records are cut to this so the consumer starts early
	};

	CommandProcessor *processor = nullptr;
	std::thread thr;
	std::vector<uint32_t> ring;
	uint64_t size = 0;

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
};
}
