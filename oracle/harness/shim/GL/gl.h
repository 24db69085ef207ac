/* Minimal stand-in for <GL/gl.h>: the reference's globalstate.h includes
 * cuda_gl_interop.h, which only needs these two typedefs to parse.  This image
 * ships no OpenGL headers.  Test/build infrastructure only. */
#pragma once
typedef unsigned int GLuint;
typedef unsigned int GLenum;
