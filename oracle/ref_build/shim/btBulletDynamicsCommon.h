// Shim for the Bullet header that the reference's include/pch.h:116 pulls in unconditionally.
// The particle path uses no physics; only the BV/BQ macro targets need to exist as types.
#pragma once
struct btVector3 { btVector3(float, float, float) {} };
struct btQuaternion { btQuaternion(float, float, float, float) {} };
// MSVC intrinsics referenced by include/pch.h:80 (DEBUG_BREAK) and src/scene/Properties.cpp.
inline void __debugbreak() { __builtin_trap(); }
inline void __nop() {}
