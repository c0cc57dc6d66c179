// Overlay stub for renderer/Text.h: Font.cpp includes it but uses nothing of it on the drawText / measureText path.
#pragma once
namespace tractor
{
class Text;
} // namespace tractor
