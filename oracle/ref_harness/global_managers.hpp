// Shim for Granite's global_managers.hpp (see context.hpp): the harness registers asset protocols it never reads here.
#pragma once
#include <memory>
#include <string>
#include "context.hpp"

namespace Granite
{
enum class PathType { File, Directory, Special };
struct FileStat { uint64_t size = 0; PathType type = PathType::File; uint64_t last_modified = 0; };

class FilesystemBackend
{
public:
	virtual ~FilesystemBackend() = default;
};

class OSFilesystem : public FilesystemBackend
{
public:
	explicit OSFilesystem(const std::string &) {}
};

class AssetManagerFilesystem : public FilesystemBackend
{
public:
	explicit AssetManagerFilesystem(const std::string &) {}
};

class Filesystem
{
public:
	void register_protocol(const std::string &, std::unique_ptr<FilesystemBackend>) {}
	bool stat(const std::string &, FileStat &) { return false; }
};

class TimelineTraceFile;
class ThreadGroup
{
public:
	TimelineTraceFile *get_timeline_trace_file() { return nullptr; }
};

namespace Global
{
inline Filesystem *filesystem() { static Filesystem fs; return &fs; }
inline ThreadGroup *thread_group() { static ThreadGroup tg; return &tg; }
inline void init() {}
inline void deinit() {}
}
}
#define GRANITE_FILESYSTEM() (::Granite::Global::filesystem())
#define GRANITE_THREAD_GROUP() (::Granite::Global::thread_group())
