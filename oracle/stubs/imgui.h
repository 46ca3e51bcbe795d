// TEST INFRASTRUCTURE ONLY.  Type-only stand-in for Dear ImGui's header, so that the reference's src/shapes.cpp -- which
// includes gui/gui.hpp for `struct Shape` -- can be compiled UNMODIFIED into oracle/_ref without the GUI toolkit (absent
// from this image).  No function, no arithmetic: only the names gui.hpp mentions.
#pragma once
struct ImVec2 {
    float x, y;
    ImVec2() : x(0.0f), y(0.0f) {}
    ImVec2(float a, float b) : x(a), y(b) {}
};
struct ImVec4 {
    float x, y, z, w;
    ImVec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    ImVec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
struct ImFont;
typedef unsigned int ImGuiID;
