// Headless stand-in for sutil/GLDisplay.h: the class the samples name, with the 5-argument display() the stock samples call AND the
// 7-argument one this fork's sutil declares (SDK/sutil/GLDisplay.h:58-64).  Never reached in batch mode (--file).
#pragma once
#include <glad/glad.h>
#include <cstdint>
#include <sutil/sutil.h>
namespace sutil {
class GLDisplay {
  public:
    GLDisplay(BufferImageFormat = sutil::BufferImageFormat::UNSIGNED_BYTE4) {}
    void display(int32_t, int32_t, int32_t, int32_t, uint32_t) const {}
    void display(int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, uint32_t) const {}
};
}
