/* Stub GL header: the reference's tsdf.cuh includes <GL/glew.h> only for render()'s
 * immediate-mode calls, which the headless emulation never executes. Test scaffolding. */
#pragma once
#define GL_TRIANGLES 4
static inline void glBegin(int) {}
static inline void glEnd() {}
static inline void glColor3f(float, float, float) {}
static inline void glVertex3f(float, float, float) {}
