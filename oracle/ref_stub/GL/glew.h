/* Typedef-only stand-in for <GL/glew.h>, used ONLY to compile the reference's
 * 3rdparty/SiftGPU/ProgramCU.cu (which includes it but needs nothing else from GL for the
 * matcher kernels) into oracle/_ref/.  Test infrastructure. */
#pragma once
typedef unsigned int GLuint;
typedef int GLint;
typedef unsigned int GLenum;
