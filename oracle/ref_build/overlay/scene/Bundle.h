// Overlay stub for scene/Bundle.h (the real one drags in the whole scene graph).  Font.cpp only touches it in
// Font::create(path, id) (Font.cpp:47-99), which the harness never calls: fonts are built with
// Font::create(family, style, size, glyphs, glyphCount, texture, format) (Font.cpp:103-161).
#pragma once
#include <string>
#include "utils/Ref.h"
namespace tractor
{
class Font;
class Bundle : public Ref
{
  public:
    static Bundle* create(const std::string&) { return nullptr; }
    std::string getObjectId(unsigned int) const { return std::string(); }
    Font* loadFont(const std::string&) { return nullptr; }
};
} // namespace tractor
