// Put this directory in front of the engine's include path and every `#include "graphics/ParticleEmitter.h"` of the
// engine and its applications (src/scene/SceneLoader.cpp:11, samples/browser/src/ParticlesSample.h) resolves to the
// B200 backend's class instead of the reference's.  Requires -DT3D_WITH_ENGINE_HEADERS (see ../../compat.h).
#pragma once
#ifndef T3D_WITH_ENGINE_HEADERS
#error "the engine include shim needs -DT3D_WITH_ENGINE_HEADERS"
#endif
#include "../../ParticleEmitter.h"
