/* euler_b200/host/scenario.h — host-side scenario handling (plain C, no CUDA).
 *
 * Keeps the reference's scenario-file format and initial-state rule byte for byte
 * (reference sim_init, main.c:209-274): 'X' solid, '0' fluid, '?' fluid + source, '=' sink,
 * anything else air; first text row -> y = ny-2, first column -> x = 1; over-long lines are
 * truncated, missing rows are air; the outer ring of cells becomes sinks; every fluid cell
 * gets 4 jittered markers drawn from xorshift64* seeded with 0x9bd185c449534b91, cells
 * visited column by column, x jitter drawn before y.  The grid size is a run-time parameter
 * here (the reference fixes X=100, Y=40 at compile time, main.c:22-25).
 */
#ifndef EULER_SCENARIO_H
#define EULER_SCENARIO_H
#include <stddef.h>
#include <stdint.h>

#define EULER_RNG_SEED 0x9bd185c449534b91ull   /* main.c:204 */

typedef struct euler_scenario {
  int nx, ny;
  uint8_t *solid, *source, *sink, *fluid;   /* [ny][nx] */
  float *markers;                           /* n_markers x (x,y) */
  size_t n_markers;
  uint64_t rng_state;                       /* randf() stream state after seeding */
} euler_scenario;

/* xorshift64* high half (misc/rng.c:5-20) and randf (main.c:203-207) */
uint32_t euler_rng_next(uint64_t *state);
float    euler_randf(uint64_t *state);

/* Parse + ring of sinks + marker seeding.  Returns 0, or -1 on allocation failure. */
int  euler_scenario_from_text(euler_scenario *s, const char *text, long length, int nx, int ny);
/* load_file (misc/file.c:5-42) + the above.  -2 if the file cannot be read. */
int  euler_scenario_load(euler_scenario *s, const char *path, int nx, int ny);
void euler_scenario_free(euler_scenario *s);
/* Re-store the seeded markers in ROW-MAJOR cell order (x fastest) instead of the reference's
 * column-major seeding order (main.c:256-257).  Positions are untouched — every marker keeps
 * the jitter it drew — only the array order changes, so consecutive markers read neighbouring
 * cells of the same grid rows (coalesced gathers on the GPU).  Only meaningful for
 * EULER_MARKERS_FAST: the reference order is observable.  0, or -1 (no memory). */
int  euler_scenario_markers_row_major(euler_scenario *s);

/* Nearest-neighbour resample of a scenario text to out_w x out_h characters (+ newlines),
 * still in the scenario format: output row j takes input row j*H/out_h, output
 * column i takes input column i*W/out_w, short input lines are padded with air.  The
 * caller frees the result with free().  *out_len receives the length. */
char *euler_scenario_resample(const char *text, long length, int out_w, int out_h, long *out_len);

/* The CURRENT state written back in the scenario-file format: text row r is grid
 * row ny-2-r, column i is x = 1+i (the inverse of main.c:220-241); 'X' solid, '?' source, '='
 * sink, '0' a cell that holds markers now (count > 0), ' ' anything else.  The outer ring of
 * sinks is implicit — sim_init adds it (main.c:244-252).  Parsing the result on the same grid
 * size reproduces the static masks exactly and makes fluid exactly the source cells plus the
 * other non-solid, non-sink cells that hold markers now; markers are re-seeded 4 per cell and
 * velocities start from rest (the format has no finer state — euler_checkpoint_save keeps all
 * of it).  The caller frees the result. */
char *euler_scenario_export(int nx, int ny, const uint8_t *solid, const uint8_t *source, const uint8_t *sink,
                            const uint8_t *count, long *out_len);

/* Synthetic scenario text for an nx x ny grid (interior (nx-2) x (ny-2) characters):
 *   "basic-fill"  walled box, fluid block resting on the floor (left 40 %, lower 50 %):
 *                 the pressure solve is active from the first sub-step
 *   "full"        walled box completely filled with fluid (worst-case traffic)
 * Returns NULL for an unknown name. */
char *euler_scenario_synthetic(const char *name, int nx, int ny, long *out_len);

#endif
