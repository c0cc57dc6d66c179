// Shim for <GL/glew.h> (include/pch.h:190). Only the scalar typedefs and the one token
// SpriteBatch.cpp tests (GL_SAMPLER_2D) are needed by the particle path.
#pragma once
typedef int GLint; typedef unsigned int GLuint; typedef unsigned int GLenum; typedef void GLvoid;
typedef int GLsizei; typedef float GLfloat; typedef unsigned char GLboolean; typedef unsigned int GLbitfield;
typedef unsigned char GLubyte; typedef char GLchar; typedef short GLshort; typedef unsigned short GLushort;
typedef double GLdouble; typedef float GLclampf; typedef long GLintptr; typedef long GLsizeiptr;
#define GL_SAMPLER_2D 0x8B5E
#define GL_NO_ERROR 0
