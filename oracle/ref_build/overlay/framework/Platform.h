// Overlay stub: FileSystem.cpp only forwards displayFileDialog to Platform (never called here).
#pragma once
#include <string>
namespace tractor
{
class Platform
{
  public:
    static std::string displayFileDialog(size_t mode, const std::string& title,
                                         const std::string& filterDescription,
                                         const std::string& filterExtensions,
                                         const std::string& initialDirectory);
};
} // namespace tractor
