/*
 * lm_oracle_contour.c -- CPU restatement of the level-set extraction the reference gets
 * from matplotlib:  plt.contour(xs, ys, Z, levels=[level])  in
 *   extract_contour     mandelbrot_boundary_sample.py:41-54
 *   extract_contour     mandelbrot_boundary_sample_spyder.py:35-43
 *
 * TEST INFRASTRUCTURE ONLY (see lm_oracle.c header).
 *
 * The arithmetic lives in a third-party dependency that is neither vendored nor pinned by
 * the reference (requirements.txt:3 "matplotlib", environment.yml:14): matplotlib calls
 * contourpy.contour_generator(name="mpl2014", corner_mask=True, chunk_size=0) and asks
 * for lines(level).  This file restates the published mpl2014 algorithm (contourpy
 * src/mpl2014.cpp: init_cache_levels, lines, get_start_edge, follow_interior, interp) for
 * the un-masked, single-chunk case, sequentially over the full quad grid with a per-quad
 * cache exactly as the original does.  contourpy/matplotlib are absent from the build
 * container and the reference ships no contour fixture => PARITY UNPINNED: this oracle
 * has been checked for internal consistency (closed loops, vertex-on-edge, level
 * separation) but not against contourpy output.
 *
 * Conventions restated:
 *   - point "above" iff z > level;
 *   - quad q = j*nx + i has corners SW=q, SE=q+1, NW=q+nx, NE=q+nx+1; edges are oriented
 *     counter-clockwise  E: SE->NE, N: NE->NW, W: NW->SW, S: SW->SE;
 *   - a vertex on edge p1->p2:  f = (z2-level)/(z2-z1);  xy = xy1*f + xy2*(1-f);
 *   - lines keep the higher side on the left; saddle quads turn right when the mean of
 *     the four corners is above the level, else left, and are visited twice;
 *   - boundary-to-boundary lines first (raster order, edges tested S,W,N,E), then interior
 *     loops in raster order of their first quad; a loop that starts on an N edge skips
 *     its initial vertex and repeats its first emitted vertex at the end.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

enum { EDGE_E = 0, EDGE_N = 1, EDGE_W = 2, EDGE_S = 3, EDGE_NONE = -1 };
enum { DIR_LEFT = 1, DIR_STRAIGHT = 0, DIR_RIGHT = -1 };

#define M_ABOVE      0x01u
#define M_VISITED    0x02u
#define M_SADDLE     0x04u
#define M_SADDLE_LEFT 0x08u
#define M_SADDLE_START_SW 0x10u

typedef struct {
    const double *x, *y, *z;
    int64_t nx, ny;
    double level;
    uint8_t* cache;          /* per point / quad */
    /* output */
    double* verts; int64_t cap_verts, n_verts;
    int64_t* offsets; int64_t cap_lines, n_lines;
    int overflow;
} Ctx;

static inline int above(const Ctx* c, int64_t p) { return (c->cache[p] & M_ABOVE) != 0; }

static inline void edge_points(const Ctx* c, int64_t quad, int edge, int64_t* p1, int64_t* p2) {
    const int64_t sw = quad, se = quad + 1, nw = quad + c->nx, ne = quad + c->nx + 1;
    switch (edge) {
        case EDGE_E: *p1 = se; *p2 = ne; break;
        case EDGE_N: *p1 = ne; *p2 = nw; break;
        case EDGE_W: *p1 = nw; *p2 = sw; break;
        default:     *p1 = sw; *p2 = se; break;
    }
}

static void push_vertex(Ctx* c, double vx, double vy) {
    if (c->n_verts < c->cap_verts) {
        c->verts[2 * c->n_verts] = vx;
        c->verts[2 * c->n_verts + 1] = vy;
    } else {
        c->overflow = 1;
    }
    c->n_verts++;
}

/* interp(point1, point2, level)   contourpy mpl2014.cpp */
static void edge_interp(Ctx* c, int64_t quad, int edge) {
    int64_t p1, p2;
    edge_points(c, quad, edge, &p1, &p2);
    const int64_t i1 = p1 % c->nx, j1 = p1 / c->nx, i2 = p2 % c->nx, j2 = p2 / c->nx;
    const double fraction = (c->z[p2] - c->level) / (c->z[p2] - c->z[p1]);
    const double vx = c->x[i1] * fraction + c->x[i2] * (1.0 - fraction);
    const double vy = c->y[j1] * fraction + c->y[j2] * (1.0 - fraction);
    push_vertex(c, vx, vy);
}

static inline int is_boundary(const Ctx* c, int64_t quad, int edge) {
    const int64_t i = quad % c->nx, j = quad / c->nx;
    switch (edge) {
        case EDGE_E: return i == c->nx - 2;
        case EDGE_N: return j == c->ny - 2;
        case EDGE_W: return i == 0;
        default:     return j == 0;
    }
}

static inline int exit_edge(int edge, int dir) {
    /* entering through `edge`; turn `dir`; returns the edge we leave through */
    switch (edge) {
        case EDGE_E: return dir == DIR_RIGHT ? EDGE_N : (dir == DIR_STRAIGHT ? EDGE_W : EDGE_S);
        case EDGE_N: return dir == DIR_RIGHT ? EDGE_W : (dir == DIR_STRAIGHT ? EDGE_S : EDGE_E);
        case EDGE_W: return dir == DIR_RIGHT ? EDGE_S : (dir == DIR_STRAIGHT ? EDGE_E : EDGE_N);
        default:     return dir == DIR_RIGHT ? EDGE_E : (dir == DIR_STRAIGHT ? EDGE_N : EDGE_W);
    }
}

static int start_edge(const Ctx* c, int64_t quad) {
    const int64_t sw = quad, se = quad + 1, nw = quad + c->nx, ne = quad + c->nx + 1;
    const unsigned config = (unsigned)(above(c, nw) << 3 | above(c, ne) << 2 | above(c, sw) << 1 | above(c, se));
    const int saddle = (c->cache[quad] & M_SADDLE) != 0;
    const int start_sw = (c->cache[quad] & M_SADDLE_START_SW) != 0;
    switch (config) {
        case 1: return EDGE_E;
        case 2: return EDGE_S;
        case 3: return EDGE_E;
        case 4: return EDGE_N;
        case 5: return EDGE_N;
        case 6: return (!saddle || start_sw) ? EDGE_S : EDGE_N;
        case 7: return EDGE_N;
        case 8: return EDGE_W;
        case 9: return (!saddle || start_sw) ? EDGE_W : EDGE_E;
        case 10: return EDGE_S;
        case 11: return EDGE_E;
        case 12: return EDGE_W;
        case 13: return EDGE_W;
        case 14: return EDGE_S;
        default: return EDGE_NONE;   /* 0 and 15 */
    }
}

/* follow_interior(...)   contourpy mpl2014.cpp (level_index 1, no corners, no parents).
 * (quad, edge) is the entry edge; on return it is the last exit edge (boundary case) or
 * the start quad-edge (closed loop).                                                      */
static void follow_interior(Ctx* c, int64_t* quad_io, int* edge_io, int want_initial_point,
                            int has_start, int64_t start_quad, int start_edge_) {
    int64_t quad = *quad_io;
    int edge = *edge_io;
    if (want_initial_point) edge_interp(c, quad, edge);
    for (;;) {
        int dir;
        if (c->cache[quad] & M_SADDLE) {
            dir = (c->cache[quad] & M_SADDLE_LEFT) ? DIR_LEFT : DIR_RIGHT;
            c->cache[quad] |= M_VISITED;
        } else {
            const int64_t sw = quad, se = quad + 1, nw = quad + c->nx, ne = quad + c->nx + 1;
            int64_t pl, pr;
            switch (edge) {
                case EDGE_E: pl = sw; pr = nw; break;
                case EDGE_N: pl = se; pr = sw; break;
                case EDGE_W: pl = ne; pr = se; break;
                default:     pl = nw; pr = ne; break;
            }
            const unsigned config = (unsigned)(above(c, pl) << 1 | above(c, pr));
            if (config == 1) {
                const double zmid = 0.25 * (c->z[sw] + c->z[se] + c->z[nw] + c->z[ne]);
                c->cache[quad] |= M_SADDLE;
                if (zmid > c->level) {
                    dir = DIR_RIGHT;
                } else {
                    dir = DIR_LEFT;
                    c->cache[quad] |= M_SADDLE_LEFT;
                }
                if (edge == EDGE_N || edge == EDGE_E) c->cache[quad] |= M_SADDLE_START_SW;
            } else {
                dir = (config == 0) ? DIR_LEFT : (config == 3 ? DIR_RIGHT : DIR_STRAIGHT);
                c->cache[quad] |= M_VISITED;
            }
        }
        edge = exit_edge(edge, dir);
        edge_interp(c, quad, edge);
        if (is_boundary(c, quad, edge)) break;
        /* move_to_next_quad */
        switch (edge) {
            case EDGE_E: quad += 1;     edge = EDGE_W; break;
            case EDGE_N: quad += c->nx; edge = EDGE_S; break;
            case EDGE_W: quad -= 1;     edge = EDGE_E; break;
            default:     quad -= c->nx; edge = EDGE_N; break;
        }
        if (has_start && quad == start_quad && edge == start_edge_) break;
    }
    *quad_io = quad;
    *edge_io = edge;
}

static void begin_line(Ctx* c) {
    if (c->n_lines < c->cap_lines) c->offsets[c->n_lines] = c->n_verts; else c->overflow = 1;
}
static void end_line(Ctx* c) {
    c->n_lines++;
    if (c->n_lines <= c->cap_lines) c->offsets[c->n_lines] = c->n_verts;
}

static int start_line(Ctx* c, int64_t quad, int edge) {
    int64_t q = quad; int e = edge;
    begin_line(c);
    follow_interior(c, &q, &e, 1, 0, 0, 0);
    end_line(c);
    return (c->cache[quad] & M_VISITED) != 0;
}

/*
 * Returns 0 on success, 1 if a capacity was too small (n_verts / n_lines then hold the
 * required sizes), -1 on allocation failure.  offsets must have room for cap_lines+1.
 */
ORACLE_API int oracle_contour_lines(const double* xs, int64_t nx, const double* ys, int64_t ny,
                                    const double* Z, double level,
                                    double* verts, int64_t cap_verts, int64_t* n_verts,
                                    int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines) {
    Ctx c;
    memset(&c, 0, sizeof(c));
    c.x = xs; c.y = ys; c.z = Z; c.nx = nx; c.ny = ny; c.level = level;
    c.verts = verts; c.cap_verts = cap_verts; c.offsets = line_offsets; c.cap_lines = cap_lines;
    *n_verts = 0; *n_lines = 0;
    if (nx < 2 || ny < 2) return 0;
    c.cache = (uint8_t*)calloc((size_t)(nx * ny), 1);
    if (!c.cache) return -1;
    for (int64_t p = 0; p < nx * ny; ++p)
        if (Z[p] > level) c.cache[p] |= M_ABOVE;

    /* lines that start and end on the boundary */
    for (int64_t j = 0; j < ny - 1; ++j) {
        for (int64_t i = 0; i < nx - 1; ++i) {
            const int64_t quad = j * nx + i;
            if (c.cache[quad] & M_VISITED) continue;
            const int64_t sw = quad, se = quad + 1, nw = quad + nx, ne = quad + nx + 1;
            if (j == 0 && above(&c, sw) && !above(&c, se) && start_line(&c, quad, EDGE_S)) continue;
            if (i == 0 && above(&c, nw) && !above(&c, sw) && start_line(&c, quad, EDGE_W)) continue;
            if (j == ny - 2 && above(&c, ne) && !above(&c, nw) && start_line(&c, quad, EDGE_N)) continue;
            if (i == nx - 2 && above(&c, se) && !above(&c, ne) && start_line(&c, quad, EDGE_E)) continue;
        }
    }
    /* interior closed loops */
    for (int64_t j = 0; j < ny - 1; ++j) {
        for (int64_t i = 0; i < nx - 1; ++i) {
            const int64_t quad = j * nx + i;
            if (c.cache[quad] & M_VISITED) continue;
            const int se = start_edge(&c, quad);
            if (se == EDGE_NONE) continue;
            int64_t q = quad; int e = se;
            const int ignore_first = (se == EDGE_N);
            begin_line(&c);
            const int64_t first_vertex = c.n_verts;
            follow_interior(&c, &q, &e, !ignore_first, 1, quad, se);
            if (ignore_first && c.n_verts > first_vertex) {
                if (first_vertex < c.cap_verts)
                    push_vertex(&c, c.verts[2 * first_vertex], c.verts[2 * first_vertex + 1]);
                else
                    push_vertex(&c, 0.0, 0.0);
            }
            end_line(&c);
            if ((c.cache[quad] & M_SADDLE) && !(c.cache[quad] & M_VISITED)) --i;   /* second pass through the saddle */
        }
    }
    free(c.cache);
    *n_verts = c.n_verts;
    *n_lines = c.n_lines;
    return c.overflow ? 1 : 0;
}
