// engine_support.cpp -- the few symbols the engine headers expect from the platform layer, for linking engine_caller.cpp
// against the B200 backend outside the engine (the headless stand-ins for Node / Scene / Texture are the overlay headers
// under oracle/ref_build/overlay; in a real engine build the engine provides all of this).
#include "pch.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "framework/FileSystem.h"
#include "framework/Platform.h"
#include "renderer/Texture.h"
#include "scene/Node.h"

GLenum __gl_error_code = 0;
ALenum __al_error_code = 0;

namespace tractor
{
void Logger::log(Level, const char* message, ...)
{
    va_list args;
    va_start(args, message);
    vfprintf(stderr, message, args);
    va_end(args);
}
void print(const char* format, ...)
{
    va_list args;
    va_start(args, format);
    vfprintf(stderr, format, args);
    va_end(args);
}
std::string Platform::displayFileDialog(size_t, const std::string&, const std::string&, const std::string&, const std::string&)
{
    return std::string();
}
// src/scene/Node.cpp:746-767: the node shares ownership of a reference-counted drawable
void Node::setDrawable(Drawable* drawable)
{
    if (_drawable == drawable) return;
    if (_drawable)
    {
        _drawable->setNode(nullptr);
        if (Ref* ref = dynamic_cast<Ref*>(_drawable)) ref->release();
    }
    _drawable = drawable;
    if (drawable)
    {
        if (Ref* ref = dynamic_cast<Ref*>(drawable)) ref->addRef();
        drawable->setNode(this);
    }
}
// only the size is ever asked for (PE.cpp:244-247): read the PNG header
Texture* Texture::create(const std::string& path, bool)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return nullptr;
    unsigned char hdr[24];
    const size_t got = fread(hdr, 1, sizeof(hdr), f);
    fclose(f);
    if (got != sizeof(hdr) || memcmp(hdr + 1, "PNG", 3) != 0) return nullptr;
    Texture* t = new Texture();
    t->_width = (hdr[16] << 24) | (hdr[17] << 16) | (hdr[18] << 8) | hdr[19];
    t->_height = (hdr[20] << 24) | (hdr[21] << 16) | (hdr[22] << 8) | hdr[23];
    return t;
}
} // namespace tractor
