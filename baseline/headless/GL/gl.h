// cuda_gl_interop.h wants the GL types; glad defines them (no system GL headers on a headless box)
#pragma once
#include <glad/glad.h>
