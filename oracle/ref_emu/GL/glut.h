/* Stub: see GL/glew.h. */
#pragma once
