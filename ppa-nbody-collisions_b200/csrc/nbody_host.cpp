// nbody_host.cpp -- the driver surface of the reference, host only (no CUDA):
//   config file parser   include/nbodyConfig.h:22-227
//   random generator     include/jbutil.h:514-562 (jbutil::randgen)
//   initial conditions   src/nbody.cu:401-416 (+ the synthetic disc / two-galaxy scenarios of BASELINE.json)
//   P5 image writer      src/nbody.cu:350-371
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

#include "nbody_b200.h"

namespace {

void emit(int fd, const std::string &s)
{
    if (fd < 0) return;
    size_t off = 0;
    while (off < s.size()) {
        ssize_t w = write(fd, s.data() + off, s.size() - off);
        if (w <= 0) return;
        off += (size_t)w;
    }
}

enum Kind { K_INT, K_FLOAT, K_STRING };
struct KeySpec {
    const char *key;      // name in the file
    const char *label;    // name in the "<label> invalid value: " message
    const char *echo;     // name in the echo line (the reference misspells two of them)
    Kind kind;
    size_t offset;
};
#define NB_OFF(f) offsetof(nb_config, f)
const KeySpec kKeys[] = {
    {"particleCount", "particleCount", "particleCount", K_INT, NB_OFF(particleCount)},                       // :36-49
    {"totalIterations", "totalIterations", "totalIterations", K_INT, NB_OFF(totalIterations)},               // :50-63
    {"save_Image_Every_Xth_Iteration", "save_Image_Every_Xth_Iteration", "save_Image_Every_Xth_Iteration", K_INT,
     NB_OFF(save_Image_Every_Xth_Iteration)},                                                                // :64-77
    {"timestep", "timestep", "timestep", K_FLOAT, NB_OFF(timestep)},                                         // :78-90
    {"minRandBodyMass", "minRandBodyMass", "minRandBodymass", K_FLOAT, NB_OFF(minRandBodyMass)},             // :91-104
    {"maxRandBodyMass", "maxRandBodyMass", "maxRandBodyMass", K_FLOAT, NB_OFF(maxRandBodyMass)},             // :105-118
    {"minRadius", "minRadius", "minRadius", K_FLOAT, NB_OFF(minRadius)},                                     // :119-132
    {"maxRadius", "maxRadius", "maxRadius", K_FLOAT, NB_OFF(maxRadius)},                                     // :133-146
    {"imgWidth", "imgWidth", "imgWidth", K_INT, NB_OFF(imgWidth)},                                           // :147-159
    {"imgHeight", "imgHeight", "imgHeight", K_INT, NB_OFF(imgHeight)},                                       // :160-172
    {"fieldWidth", "fieldWidth", "fieldWidth", K_INT, NB_OFF(fieldWidth)},                                   // :173-186
    {"fieldHeight", "fieldHeight", "fieldHeight", K_INT, NB_OFF(fieldHeight)},                               // :187-200
    {"imagePath", "imagePath", "imagePath", K_STRING, NB_OFF(imagePath)},                                    // :201-207
    {"radiusGrowthRate", "growthRate", "growthRate", K_FLOAT, NB_OFF(growthRate)},                           // :208-221
};

}  // namespace

extern "C" {

int nb_config_parse(const char *path, nb_config *cfg, int echo_fd)
{
    if (!path || !cfg) return NB_ERR_INVALID;
    std::ifstream in(path);
    if (!in.is_open()) {
        emit(echo_fd, "Error opening config file! Exiting...\n");                                            // :26
        return NB_ERR_IO;
    }
    std::string line;
    while (std::getline(in, line)) {
        const size_t delim = line.find("=");
        const std::string name = line.substr(0, delim);
        // substr(npos + 1) == substr(0): a line without '=' hands the whole line to stoi/stof (:40)
        const std::string value = line.substr(delim + 1);
        const KeySpec *spec = nullptr;
        for (const KeySpec &k : kKeys)
            if (name == k.key) {
                spec = &k;
                break;
            }
        if (!spec) {
            emit(echo_fd, "Invalid variable: " + name + "\n");                                               // :222-224
            continue;
        }
        std::ostringstream os;                    // default ostream formatting, like std::cout
        char *field = reinterpret_cast<char *>(cfg) + spec->offset;
        try {
            if (spec->kind == K_INT) {
                const int v = std::stoi(value);
                os << spec->echo << "=" << v << "\n";
                memcpy(field, &v, sizeof(int));
            } else if (spec->kind == K_FLOAT) {
                const float v = std::stof(value);  // a trailing 'f' ("0.2f") is ignored by stof
                os << spec->echo << "=" << v << "\n";
                memcpy(field, &v, sizeof(float));
            } else {
                os << spec->echo << "=" << value << "\n";
                snprintf(field, sizeof(cfg->imagePath), "%s", value.c_str());
            }
        } catch (std::exception const &e) {
            emit(echo_fd, std::string(spec->label) + " invalid value: " + e.what() + "\n");
            return NB_ERR_INVALID;                 // the reference calls exit(1) here
        }
        emit(echo_fd, os.str());
    }
    return NB_OK;
}

// ---- jbutil::randgen (Numerical Recipes "Ran"), include/jbutil.h:514-562 ----------------------
static inline void rng_advance(nb_rng *g)                                       // jbutil.h:537-544
{
    g->u = g->u * 2862933555777941757ULL + 7046029254386353087ULL;
    g->v ^= g->v >> 17;
    g->v ^= g->v << 31;
    g->v ^= g->v >> 8;
    g->w = 4294957665ULL * (g->w & 0xffffffffULL) + (g->w >> 32);
}

uint64_t nb_rng_ival64(nb_rng *g)                                               // jbutil.h:546-553
{
    rng_advance(g);
    uint64_t x = g->u ^ (g->u << 21);
    x ^= x >> 35;
    x ^= x << 4;
    return (x + g->v) ^ g->w;
}

void nb_rng_seed(nb_rng *g, uint64_t seed)                                      // jbutil.h:525-535
{
    g->v = 4101842887655102017ULL;
    g->w = 1;
    g->u = seed ^ g->v;
    nb_rng_ival64(g);
    g->v = g->u;
    nb_rng_ival64(g);
    g->w = g->v;
    nb_rng_ival64(g);
}

double nb_rng_fval(nb_rng *g) { return 5.42101086242752217E-20 * (double)nb_rng_ival64(g); }    // jbutil.h:554-557

double nb_rng_fval_range(nb_rng *g, double a, double b) { return nb_rng_fval(g) * (b - a) + a; } // jbutil.h:558-561

// ---- initial conditions ------------------------------------------------------------------------
// Fills `count` bodies starting at index `first` of an n-body BodiesData block with a uniform disc:
// per body four draws in the order u1, u2, m, r (the reference's draw order with the position pair
// reinterpreted as radius^2 fraction and angle), computed in double and stored as float.
static void fill_disc(float *block, int n, int first, int count, uint64_t seed, double cx, double cy, double R,
                      double bulk_vx, double bulk_vy, double omega, const nb_scenario *sc)
{
    float *pos = block, *vel = block + 2 * (size_t)n, *mass = block + 4 * (size_t)n, *rad = block + 5 * (size_t)n;
    nb_rng g;
    nb_rng_seed(&g, seed);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < count; ++k) {
        const int b = first + k;
        const double u1 = nb_rng_fval(&g), u2 = nb_rng_fval(&g);
        const double rr = R * std::sqrt(u1), th = two_pi * u2;
        const double x = rr * std::cos(th), y = rr * std::sin(th);
        pos[2 * b] = (float)(cx + x);
        pos[2 * b + 1] = (float)(cy + y);
        vel[2 * b] = (float)(bulk_vx - omega * y);
        vel[2 * b + 1] = (float)(bulk_vy + omega * x);
        mass[b] = (float)nb_rng_fval_range(&g, sc->min_mass, sc->max_mass);
        rad[b] = (float)nb_rng_fval_range(&g, sc->min_radius, sc->max_radius);
    }
}

int nb_generate(const nb_scenario *sc, void *bodies)
{
    if (!sc || !bodies || sc->n < 0) return NB_ERR_INVALID;
    float *block = static_cast<float *>(bodies);
    const int n = sc->n;
    if (sc->kind == NB_SCENARIO_SQUARE) {
        // src/nbody.cu:401-416: four draws per body in the order x, y, m, r; float = double - int; v = 0
        float *pos = block, *vel = block + 2 * (size_t)n, *mass = block + 4 * (size_t)n, *rad = block + 5 * (size_t)n;
        nb_rng g;
        nb_rng_seed(&g, sc->seed);                                               // :403
        const int dw = sc->field_w << 1, dh = sc->field_h << 1;                  // :388,390
        for (int b = 0; b < n; ++b) {
            pos[2 * b] = (float)(nb_rng_fval_range(&g, 0, dw) - sc->field_w);
            pos[2 * b + 1] = (float)(nb_rng_fval_range(&g, 0, dh) - sc->field_h);
            mass[b] = (float)nb_rng_fval_range(&g, sc->min_mass, sc->max_mass);
            rad[b] = (float)nb_rng_fval_range(&g, sc->min_radius, sc->max_radius);
            vel[2 * b] = 0.f;
            vel[2 * b + 1] = 0.f;
        }
        return NB_OK;
    }
    if (sc->kind == NB_SCENARIO_DISC) {
        if (!(sc->extent > 0)) return NB_ERR_INVALID;
        fill_disc(block, n, 0, n, sc->seed, 0.0, 0.0, sc->extent, 0.0, 0.0, 0.0, sc);
        return NB_OK;
    }
    if (sc->kind == NB_SCENARIO_TWO_GALAXY) {
        // two counter-rotating discs of radius R on an encounter course:
        // centres (-+1.5 R, -+0.25 R), bulk velocities (+-400, 0), solid-body spin +-2e-4 rad per unit time
        if (!(sc->extent > 0)) return NB_ERR_INVALID;
        const double R = sc->extent;
        const int n0 = n / 2;
        fill_disc(block, n, 0, n0, sc->seed, -1.5 * R, -0.25 * R, R, 400.0, 0.0, 2.0e-4, sc);
        fill_disc(block, n, n0, n - n0, sc->seed + 1, 1.5 * R, 0.25 * R, R, -400.0, 0.0, -2.0e-4, sc);
        return NB_OK;
    }
    return NB_ERR_INVALID;
}

// ---- image writer, src/nbody.cu:350-371 (one buffered write instead of one << per byte) ----------
int nb_write_pgm(const char *path, const uint8_t *image, int w, int h)
{
    if (!path || !image || w <= 0 || h <= 0) return NB_ERR_INVALID;
    FILE *f = fopen(path, "wb");
    if (!f) return NB_ERR_IO;                      // the reference prints "Error writing image to file" and exits
    fprintf(f, "P5\n%d %d\n255\n", w, h);
    const size_t bytes = (size_t)w * h;
    const size_t wr = fwrite(image, 1, bytes, f);
    const int rc = fclose(f);
    return (wr == bytes && rc == 0) ? NB_OK : NB_ERR_IO;
}

}  // extern "C"
