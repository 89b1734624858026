/* euler_b200/host/render.c — see render.h. */
#define _POSIX_C_SOURCE 200809L
#include "render.h"

#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/ioctl.h>
#include <termios.h>
#include <time.h>
#include <unistd.h>

static struct termios g_saved;
static int g_raw = 0;

int euler_tty_window_size(int *rows, int *cols) {
  struct winsize ws;
  if (ioctl(STDOUT_FILENO, TIOCGWINSZ, &ws) == -1 || ws.ws_col == 0) return -1;
  *rows = ws.ws_row; *cols = ws.ws_col;
  return 0;
}

void euler_tty_restore(void) {
  if (g_raw) { tcsetattr(STDIN_FILENO, TCSAFLUSH, &g_saved); g_raw = 0; }
  const char show[] = "\x1b[?25h\x1b[0m";
  if (write(STDOUT_FILENO, show, sizeof show - 1) < 0) {}
}

int euler_tty_raw_mode(void) {
  if (tcgetattr(STDIN_FILENO, &g_saved) == -1) return -1;
  struct termios t = g_saved;
  t.c_iflag &= ~(unsigned)(BRKINT | ICRNL | INPCK | ISTRIP | IXON);
  t.c_oflag &= ~(unsigned)OPOST;
  t.c_cflag |= CS8;
  t.c_lflag &= ~(unsigned)(ECHO | ICANON | IEXTEN | ISIG);
  t.c_cc[VMIN] = 0; t.c_cc[VTIME] = 0;
  if (tcsetattr(STDIN_FILENO, TCSAFLUSH, &t) == -1) return -1;
  g_raw = 1;
  atexit(euler_tty_restore);
  return 0;
}

void euler_tty_clear(void) {
  const char seq[] = "\x1b[2J\x1b[H";
  if (write(STDOUT_FILENO, seq, sizeof seq - 1) < 0) {}
}

char euler_tty_read_key(void) {
  char c = 0;
  if (read(STDIN_FILENO, &c, 1) == -1 && errno != EAGAIN && errno != EINTR) return 'q';
  return c;
}

static void put(euler_screen *s, const char *p, size_t n) {
  if (s->len + n + 1 > s->cap) {
    size_t cap = s->cap ? s->cap * 2 : 1 << 16;
    while (cap < s->len + n + 1) cap *= 2;
    char *nb = realloc(s->buf, cap);
    if (!nb) return;
    s->buf = nb; s->cap = cap;
  }
  memcpy(s->buf + s->len, p, n);
  s->len += n;
}
#define PUTS(s, lit) put(s, lit, sizeof(lit) - 1)

void euler_draw(euler_screen *s, int nx, int ny, const uint8_t *solid, const uint8_t *sink,
                const uint8_t *count) {
  static const char glyph[4] = {' ', 'o', 'O', '0'};
  s->len = 0;
  PUTS(s, "\x1b[H");
  int y_low = ny - 1 - s->rows;
  if (y_low < 1) y_low = 1;
  for (int y = ny - 2; y >= y_low; --y) {
    int wet = 0;
    for (int x = 1; x < nx - 1 && x < s->cols + 1; ++x) {
      const size_t c = (size_t)y * nx + x;
      if (solid[c]) { if (wet) PUTS(s, "\x1b[0m"); PUTS(s, "X"); wet = 0; }
      /* like main.c:928-932 a sink resets the colour but NOT prev_water: water right after a sink
       * is drawn without a new blue escape.  Kept: the bytes equal the reference's. */
      else if (sink[c]) { if (wet) PUTS(s, "\x1b[0m"); PUTS(s, "="); }
      else {
        const int level = count[c] < 3 ? count[c] : 3;
        if (level && !wet) PUTS(s, "\x1b[34m");
        else if (!level && wet) PUTS(s, "\x1b[0m");
        put(s, &glyph[level], 1);
        wet = level != 0;
      }
    }
    PUTS(s, "\x1b[0m\x1b[K");
    if (y > y_low) PUTS(s, "\r\n");
  }
  PUTS(s, "\x1b[?25l");
  if (write(STDOUT_FILENO, s->buf, s->len) < 0) {}
}

int euler_color_byte(float linear) {
  /* misc/color.h: float_to_byte_color(linear_to_sRGB(x)) */
  const float end = nextafterf(256.f, 0.f);
  float v = end * powf(linear, 1 / 2.2f);
  if (!(v > 0.f)) v = 0.f;
  if (v > end) v = end;
  return (int)v;
}

static void draw_frame(euler_screen *s, int nx, int ny, const uint8_t *solid, const uint8_t *sink,
                       const uint8_t *count, const float *r, const float *g, const float *b) {
  static const char glyph[4] = {' ', 'o', 'O', '0'};
  s->len = 0;
  PUTS(s, "\x1b[H");
  int y_low = ny - 1 - s->rows;
  if (y_low < 1) y_low = 1;
  for (int y = ny - 2; y >= y_low; --y) {
    int wet = 0;
    for (int x = 1; x < nx - 1 && x < s->cols + 1; ++x) {
      const size_t c = (size_t)y * nx + x;
      if (solid[c]) { if (wet) PUTS(s, "\x1b[0m"); PUTS(s, "X"); wet = 0; }
      else if (sink[c]) { if (wet) PUTS(s, "\x1b[0m"); PUTS(s, "="); }   /* prev_water kept, main.c:928-932 */
      else {
        const int level = count[c] < 3 ? count[c] : 3;
        if (level) {
          char esc[24];
          int n = snprintf(esc, sizeof esc, "\x1b[38;2;%d;%d;%dm", euler_color_byte(r[c]),
                           euler_color_byte(g[c]), euler_color_byte(b[c]));
          if (n > 0 && n < (int)sizeof esc) put(s, esc, (size_t)n);
        } else if (wet) PUTS(s, "\x1b[0m");
        put(s, &glyph[level], 1);
        wet = level != 0;
      }
    }
    PUTS(s, "\x1b[0m\x1b[K");
    if (y > y_low) PUTS(s, "\r\n");
  }
  PUTS(s, "\x1b[?25l");
  if (write(STDOUT_FILENO, s->buf, s->len) < 0) {}
}

void euler_draw_rainbow(euler_screen *s, int nx, int ny, const uint8_t *solid, const uint8_t *sink,
                        const uint8_t *count, const float *r, const float *g, const float *b) {
  draw_frame(s, nx, ny, solid, sink, count, r, g, b);
}

size_t euler_draw_plain(char *dst, size_t cap, int nx, int ny, int max_cols, int max_rows,
                        const uint8_t *solid, const uint8_t *sink, const uint8_t *count) {
  static const char glyph[4] = {' ', 'o', 'O', '0'};
  size_t n = 0;
  int y_low = ny - 1 - max_rows;
  if (y_low < 1) y_low = 1;
  for (int y = ny - 2; y >= y_low; --y) {
    for (int x = 1; x < nx - 1 && x < max_cols + 1; ++x) {
      const size_t c = (size_t)y * nx + x;
      char ch = solid[c] ? 'X' : sink[c] ? '=' : glyph[count[c] < 3 ? count[c] : 3];
      if (n + 2 < cap) dst[n++] = ch;
    }
    if (n + 2 < cap) dst[n++] = '\n';
  }
  if (cap) dst[n < cap ? n : cap - 1] = '\0';
  return n;
}

void euler_screen_free(euler_screen *s) { free(s->buf); memset(s, 0, sizeof *s); }

euler_time euler_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  euler_time t = { (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec };
  return t;
}

euler_time euler_wait_until(euler_time start, long long period_ns) {
  euler_time target = { start.ns + period_ns };
  euler_time now = euler_now();
  if (now.ns < target.ns) {
    struct timespec d = { (time_t)((target.ns - now.ns) / 1000000000ll), (long)((target.ns - now.ns) % 1000000000ll) };
    nanosleep(&d, NULL);
    return target;
  }
  return now;       /* running late: do not try to catch up */
}
