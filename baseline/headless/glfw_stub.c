/* Headless stand-in for the GLFW library: the reference samples link against GLFW even when they render to a file
 * (--file out.ppm --no-gl-interop never reaches a window call: SDK/optixPathTracer/optixPathTracer.cpp:1051-1085).  Every entry point the
 * samples, sutil and the vendored imgui back end reference is defined here and fails the way a machine without a display would. */
#define GLFW_INCLUDE_NONE
#include <GLFW/glfw3.h>
#include <stddef.h>

int glfwInit(void) { return GLFW_FALSE; }
void glfwTerminate(void) {}
GLFWerrorfun glfwSetErrorCallback(GLFWerrorfun f) { (void)f; return NULL; }
void glfwWindowHint(int hint, int value) { (void)hint; (void)value; }
GLFWwindow* glfwCreateWindow(int w, int h, const char* t, GLFWmonitor* m, GLFWwindow* s) { (void)w; (void)h; (void)t; (void)m; (void)s; return NULL; }
void glfwDestroyWindow(GLFWwindow* w) { (void)w; }
int glfwWindowShouldClose(GLFWwindow* w) { (void)w; return GLFW_TRUE; }
void glfwSetWindowShouldClose(GLFWwindow* w, int v) { (void)w; (void)v; }
void glfwMakeContextCurrent(GLFWwindow* w) { (void)w; }
void glfwSwapInterval(int i) { (void)i; }
void glfwSwapBuffers(GLFWwindow* w) { (void)w; }
void glfwPollEvents(void) {}
void glfwWaitEvents(void) {}
void glfwGetFramebufferSize(GLFWwindow* w, int* x, int* y) { (void)w; if (x) *x = 0; if (y) *y = 0; }
void glfwGetWindowSize(GLFWwindow* w, int* x, int* y) { (void)w; if (x) *x = 0; if (y) *y = 0; }
int glfwGetWindowAttrib(GLFWwindow* w, int a) { (void)w; (void)a; return 0; }
void glfwSetWindowUserPointer(GLFWwindow* w, void* p) { (void)w; (void)p; }
void* glfwGetWindowUserPointer(GLFWwindow* w) { (void)w; return NULL; }
GLFWkeyfun glfwSetKeyCallback(GLFWwindow* w, GLFWkeyfun f) { (void)w; (void)f; return NULL; }
GLFWcharfun glfwSetCharCallback(GLFWwindow* w, GLFWcharfun f) { (void)w; (void)f; return NULL; }
GLFWscrollfun glfwSetScrollCallback(GLFWwindow* w, GLFWscrollfun f) { (void)w; (void)f; return NULL; }
GLFWmousebuttonfun glfwSetMouseButtonCallback(GLFWwindow* w, GLFWmousebuttonfun f) { (void)w; (void)f; return NULL; }
GLFWcursorposfun glfwSetCursorPosCallback(GLFWwindow* w, GLFWcursorposfun f) { (void)w; (void)f; return NULL; }
GLFWwindowsizefun glfwSetWindowSizeCallback(GLFWwindow* w, GLFWwindowsizefun f) { (void)w; (void)f; return NULL; }
GLFWwindowiconifyfun glfwSetWindowIconifyCallback(GLFWwindow* w, GLFWwindowiconifyfun f) { (void)w; (void)f; return NULL; }
void glfwSetInputMode(GLFWwindow* w, int m, int v) { (void)w; (void)m; (void)v; }
int glfwGetInputMode(GLFWwindow* w, int m) { (void)w; (void)m; return 0; }
void glfwGetCursorPos(GLFWwindow* w, double* x, double* y) { (void)w; if (x) *x = 0; if (y) *y = 0; }
void glfwSetCursorPos(GLFWwindow* w, double x, double y) { (void)w; (void)x; (void)y; }
int glfwGetMouseButton(GLFWwindow* w, int b) { (void)w; (void)b; return 0; }
GLFWcursor* glfwCreateStandardCursor(int shape) { (void)shape; return NULL; }
void glfwDestroyCursor(GLFWcursor* c) { (void)c; }
void glfwSetCursor(GLFWwindow* w, GLFWcursor* c) { (void)w; (void)c; }
void glfwSetClipboardString(GLFWwindow* w, const char* s) { (void)w; (void)s; }
const char* glfwGetClipboardString(GLFWwindow* w) { (void)w; return ""; }
double glfwGetTime(void) { return 0.0; }
const unsigned char* glfwGetJoystickButtons(int jid, int* count) { (void)jid; if (count) *count = 0; return NULL; }
const float* glfwGetJoystickAxes(int jid, int* count) { (void)jid; if (count) *count = 0; return NULL; }
