// boost/filesystem.hpp — what dqn_main.cpp:7,:15,:233,:405 uses of it (path, path::native(),
// is_regular_file) is std::filesystem, name for name.
#pragma once
#include <filesystem>
namespace boost {
namespace filesystem {
using std::filesystem::path;
using std::filesystem::is_regular_file;
using std::filesystem::exists;
using std::filesystem::create_directories;
}  // namespace filesystem
}  // namespace boost
