/* euler_b200/host/render.h — ASCII rendering and tty handling on the host (off the timed path).
 * Same picture as the reference's draw_rows (main.c:914-951): marker count -> " oO0", solid
 * 'X', sink '=', water in blue, one text row per grid row from the top, clipped to the
 * terminal window. */
#ifndef EULER_RENDER_H
#define EULER_RENDER_H
#include <stddef.h>
#include <stdint.h>

typedef struct euler_screen {
  char *buf;
  size_t len, cap;
  int cols, rows;            /* terminal window */
} euler_screen;

int  euler_tty_window_size(int *rows, int *cols);     /* -1 if stdout is not a terminal */
int  euler_tty_raw_mode(void);                        /* restored automatically at exit */
void euler_tty_restore(void);
void euler_tty_clear(void);
/* Non-blocking single key read; 0 if none. */
char euler_tty_read_key(void);

/* Compose one frame into scr->buf (cursor home, rows, hide cursor) and write it to stdout. */
void euler_draw(euler_screen *scr, int nx, int ny, const uint8_t *solid, const uint8_t *sink,
                const uint8_t *marker_count);
/* Same with --rainbow: every wet cell in its own 24-bit colour, linear RGB -> sRGB bytes like
 * buffer_append_color (main.c:902-912, misc/color.h).  r, g, b are [ny][nx] planes. */
void euler_draw_rainbow(euler_screen *scr, int nx, int ny, const uint8_t *solid, const uint8_t *sink,
                        const uint8_t *marker_count, const float *r, const float *g, const float *b);
/* misc/color.h: linear [0,1] -> sRGB byte (x^(1/2.2) approximation, clamped) */
int euler_color_byte(float linear);
/* Same picture without escape codes into a caller buffer (for --headless --print). Returns
 * the number of bytes written (excluding the terminating NUL). */
size_t euler_draw_plain(char *dst, size_t cap, int nx, int ny, int max_cols, int max_rows,
                        const uint8_t *solid, const uint8_t *sink, const uint8_t *marker_count);
void euler_screen_free(euler_screen *scr);

/* Sleep until `start + period_ns`, return the new start (10 fps pacing, main.c:1036). */
typedef struct euler_time { long long ns; } euler_time;
euler_time euler_now(void);
euler_time euler_wait_until(euler_time start, long long period_ns);
#endif
