// TEST INFRASTRUCTURE ONLY: type-only stand-in (see ../imgui.h).
#pragma once
struct GLFWwindow;
