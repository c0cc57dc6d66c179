// Overlay stub: ParticleEmitter.h includes math/Transform.h only to reach Matrix/Quaternion/Vector3.
#pragma once
#include "math/Matrix.h"
#include "math/Quaternion.h"
#include "math/Vector3.h"
