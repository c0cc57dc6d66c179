// Case alias: include/animation/AnimationClip.h:20 spells the header "utils/ref.h" (fine on
// Windows, missing on a case-sensitive filesystem).
#pragma once
#include "utils/Ref.h"
