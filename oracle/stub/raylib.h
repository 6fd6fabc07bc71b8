/* Minimal stand-in for raylib.h -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference env headers (pufferlib/ocean/drone_race/dronelib.h:13,
 * drone_race.h:13, drone_swarm.h:13) include "raylib.h" for their renderer
 * (c_render).  raylib is a build-time download of the reference's setup.py and
 * is not vendored, so the oracle build supplies the handful of public raylib
 * types/constants the renderer mentions and turns every draw/window call into
 * a no-op.  None of this is on the env-step path: it only lets the unmodified
 * reference sources compile.  PI keeps raylib's float literal because
 * drone_swarm.h:249 uses it arithmetically.
 */
#ifndef B2D_ORACLE_RAYLIB_STUB_H
#define B2D_ORACLE_RAYLIB_STUB_H
#include <stdbool.h>
#include <stdarg.h>

#ifndef PI
#define PI 3.14159265358979323846f
#endif

typedef struct Vector2 { float x, y; } Vector2;
typedef struct Vector3 { float x, y, z; } Vector3;
typedef struct Color { unsigned char r, g, b, a; } Color;
typedef struct Camera3D {
    Vector3 position, target, up;
    float fovy;
    int projection;
} Camera3D;
typedef Camera3D Camera;

#define B2D_RGBA(r_, g_, b_) ((Color){(r_), (g_), (b_), 255})
#define LIGHTGRAY B2D_RGBA(200, 200, 200)
#define GRAY      B2D_RGBA(130, 130, 130)
#define DARKGRAY  B2D_RGBA(80, 80, 80)
#define YELLOW    B2D_RGBA(253, 249, 0)
#define GOLD      B2D_RGBA(255, 203, 0)
#define ORANGE    B2D_RGBA(255, 161, 0)
#define PINK      B2D_RGBA(255, 109, 194)
#define RED       B2D_RGBA(230, 41, 55)
#define MAROON    B2D_RGBA(190, 33, 55)
#define GREEN     B2D_RGBA(0, 228, 48)
#define LIME      B2D_RGBA(0, 158, 47)
#define DARKGREEN B2D_RGBA(0, 117, 44)
#define SKYBLUE   B2D_RGBA(102, 191, 255)
#define BLUE      B2D_RGBA(0, 121, 241)
#define DARKBLUE  B2D_RGBA(0, 82, 172)
#define PURPLE    B2D_RGBA(200, 122, 255)
#define VIOLET    B2D_RGBA(135, 60, 190)
#define MAGENTA   B2D_RGBA(255, 0, 255)
#define WHITE     B2D_RGBA(255, 255, 255)
#define BLACK     B2D_RGBA(0, 0, 0)
#define RAYWHITE  B2D_RGBA(245, 245, 245)

enum { CAMERA_PERSPECTIVE = 0, CAMERA_ORTHOGRAPHIC = 1 };
enum { MOUSE_BUTTON_LEFT = 0, MOUSE_BUTTON_RIGHT = 1 };
enum { KEY_ESCAPE = 256, KEY_SPACE = 32, KEY_TAB = 258, KEY_LEFT_SHIFT = 340 };
enum { FLAG_MSAA_4X_HINT = 0x20 };
enum { LOG_INFO = 3, LOG_WARNING = 4, LOG_ERROR = 5 };

#define B2D_NOP static inline __attribute__((unused))
B2D_NOP void InitWindow(int w, int h, const char *t) { (void)w; (void)h; (void)t; }
B2D_NOP void CloseWindow(void) {}
B2D_NOP bool WindowShouldClose(void) { return false; }
B2D_NOP bool IsWindowReady(void) { return false; }
B2D_NOP void SetTargetFPS(int f) { (void)f; }
B2D_NOP void SetConfigFlags(unsigned int f) { (void)f; }
B2D_NOP void TraceLog(int lvl, const char *fmt, ...) { (void)lvl; (void)fmt; }
B2D_NOP const char *TextFormat(const char *fmt, ...) { return fmt; }
B2D_NOP bool IsKeyDown(int k) { (void)k; return false; }
B2D_NOP bool IsKeyPressed(int k) { (void)k; return false; }
B2D_NOP bool IsMouseButtonPressed(int b) { (void)b; return false; }
B2D_NOP bool IsMouseButtonReleased(int b) { (void)b; return false; }
B2D_NOP bool IsMouseButtonDown(int b) { (void)b; return false; }
B2D_NOP Vector2 GetMousePosition(void) { return (Vector2){0.0f, 0.0f}; }
B2D_NOP float GetMouseWheelMove(void) { return 0.0f; }
B2D_NOP void BeginDrawing(void) {}
B2D_NOP void EndDrawing(void) {}
B2D_NOP void ClearBackground(Color c) { (void)c; }
B2D_NOP void BeginMode3D(Camera3D c) { (void)c; }
B2D_NOP void EndMode3D(void) {}
B2D_NOP Color ColorAlpha(Color c, float a) { (void)a; return c; }
B2D_NOP void DrawText(const char *t, int x, int y, int s, Color c) { (void)t; (void)x; (void)y; (void)s; (void)c; }
B2D_NOP void DrawSphere(Vector3 p, float r, Color c) { (void)p; (void)r; (void)c; }
B2D_NOP void DrawLine3D(Vector3 a, Vector3 b, Color c) { (void)a; (void)b; (void)c; }
B2D_NOP void DrawCubeWires(Vector3 p, float w, float h, float l, Color c) { (void)p; (void)w; (void)h; (void)l; (void)c; }
B2D_NOP void DrawCylinderEx(Vector3 a, Vector3 b, float r0, float r1, int s, Color c) { (void)a; (void)b; (void)r0; (void)r1; (void)s; (void)c; }
B2D_NOP void DrawCylinderWiresEx(Vector3 a, Vector3 b, float r0, float r1, int s, Color c) { (void)a; (void)b; (void)r0; (void)r1; (void)s; (void)c; }
#endif
