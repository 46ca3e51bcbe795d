// TEST INFRASTRUCTURE ONLY: type-only stand-in (see ../imgui.h).
#pragma once
typedef unsigned int GLuint;
