#pragma once
typedef int ALenum; typedef unsigned int ALuint; typedef int ALint; typedef float ALfloat;
