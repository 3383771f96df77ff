#!/usr/bin/env python
"""Generate tests/golden/host_golden.json from the REFERENCE'S OWN HEADERS.

Run in the build container (needs /root/reference).  A throw-away C++ program
that #includes the reference's jbutil.h / nbodyConfig.h / vec2f.h where they
lie is compiled into /tmp and run; nothing from the reference is copied into
the repo.  The JSON pins:
  * jbutil::randgen after seed(1024): raw 64-bit stream and fval() doubles
    (include/jbutil.h:514-562);
  * the initial bodies of the shipped scenario exactly as src/nbody.cu:401-416
    generates them (bit patterns of the first/last bodies, FNV-1a-64 of the
    whole BodiesData block, double sums);
  * parseConfigFile()'s stdout echo and parsed values for the shipped
    nbodyConfig.txt and for a file with quirks (include/nbodyConfig.h:22-227);
  * a few Vec2f operator results (include/vec2f.h:44-98).
"""
import json
import subprocess
import sys
import tempfile
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "host_golden.json"

CPP = r'''
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "jbutil.h"
#include "vec2f.h"
#include "nbodyConfig.h"

static uint64_t fnv(const void* d, size_t n) {
    const unsigned char* p = (const unsigned char*)d; uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ULL; } return h;
}
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv) {
    std::string mode = argv[1];
    if (mode == "rng") {
        jbutil::randgen g; g.seed(1024);
        for (int i = 0; i < 8; ++i) printf("%llu\n", (unsigned long long)g.ival64());
        jbutil::randgen h; h.seed(1024);
        for (int i = 0; i < 8; ++i) printf("%.17g\n", h.fval());
        jbutil::randgen k; k.seed(7);
        for (int i = 0; i < 4; ++i) printf("%llu\n", (unsigned long long)k.ival64());
    } else if (mode == "config") {
        ConfigData c = parseConfigFile(argv[2]);
        printf("@@ %d %d %d %.9g %.9g %.9g %.9g %.9g %.9g %d %d %d %d [%s]\n", c.particleCount, c.totalIterations,
               c.save_Image_Every_Xth_Iteration, c.timestep, c.minRandBodyMass, c.maxRandBodyMass,
               c.minRadius, c.maxRadius, c.growthRate, c.imgWidth, c.imgHeight, c.fieldWidth, c.fieldHeight,
               c.imagePath.c_str());
    } else if (mode == "init") {
        // src/nbody.cu:381-416, same statements, same types.
        ConfigData config = parseConfigFile(argv[2]);
        const int particleCount = config.particleCount;
        const float minBodyMass = config.minRandBodyMass;
        const float maxBodyMass = config.maxRandBodyMass;
        int fieldWidth = config.fieldWidth, doubleFieldWidth = fieldWidth << 1;
        int fieldHeight = config.fieldHeight, doubleFieldHeight = fieldHeight << 1;
        std::vector<float> block(6 * (size_t)particleCount);
        Vec2f* P = (Vec2f*)block.data(); Vec2f* V = P + particleCount;
        float* M = (float*)(V + particleCount); float* R = M + particleCount;
        jbutil::randgen gen; gen.seed(1024);
        float x, y, m, r;
        double sx = 0, sy = 0, sm = 0, sr = 0;
        for (int b = 0; b < particleCount; ++b) {
            x = gen.fval(0, doubleFieldWidth) - fieldWidth;
            y = gen.fval(0, doubleFieldHeight) - fieldHeight;
            m = gen.fval(minBodyMass, maxBodyMass);
            r = gen.fval(config.minRadius, config.maxRadius);
            P[b] = Vec2f(x, y); V[b] = Vec2f(0.f, 0.f); M[b] = m; R[b] = r;
            sx += x; sy += y; sm += m; sr += r;
        }
        printf("@@ %d %016llx %.17g %.17g %.17g %.17g\n", particleCount,
               (unsigned long long)fnv(block.data(), block.size() * 4), sx, sy, sm, sr);
        int idx[6] = {0, 1, 2, 3, particleCount / 2, particleCount - 1};
        for (int q = 0; q < 6; ++q) { int b = idx[q];
            printf("## %d %08x %08x %08x %08x\n", b, bits(P[b].X), bits(P[b].Y), bits(M[b]), bits(R[b])); }
    } else if (mode == "vec2f") {
        Vec2f a(3.f, -4.f), b(0.5f, 7.f);
        Vec2f c = a * 3.f; Vec2f d = a / 3.f; Vec2f e = a * b; Vec2f f = a + b; Vec2f g = a - b; Vec2f h = -a;
        Vec2f s(2.5f); Vec2f k = 2.f * b;
        printf("%08x %08x\n", bits(c.X), bits(c.Y)); printf("%08x %08x\n", bits(d.X), bits(d.Y));
        printf("%08x %08x\n", bits(e.X), bits(e.Y)); printf("%08x %08x\n", bits(f.X), bits(f.Y));
        printf("%08x %08x\n", bits(g.X), bits(g.Y)); printf("%08x %08x\n", bits(h.X), bits(h.Y));
        printf("%08x %08x\n", bits(s.X), bits(s.Y)); printf("%08x %08x\n", bits(k.X), bits(k.Y));
        printf("%08x\n", bits(a.length())); printf("%08x\n", bits(Vec2f(1e-3f, 7.f).length()));
        printf("%zu\n", sizeof(Vec2f));
    }
    return 0;
}
'''

QUIRKY = """particleCount=300
bogusKey=12

timestep=0.05f
radiusGrowthRate=0.25
minRandBodyMass=1e4f
maxRandBodyMass=1e17f
minRadius=50.f
maxRadius=200.f
totalIterations=7
save_Image_Every_Xth_Iteration=3
imgWidth=64
imgHeight=32
fieldWidth=2000
fieldHeight=1000
imagePath=some dir/with=equals
"""


def run(exe, *args):
    return subprocess.run([str(exe), *args], check=True, capture_output=True, text=True).stdout


def main():
    if not (REF / "include" / "jbutil.h").exists():
        sys.exit("reference not present; run this in the build container")
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "g.cpp").write_text(CPP)
        exe = td / "g"
        subprocess.run(["/usr/bin/g++", "-O1", "-ffp-contract=off", f"-I{REF / 'include'}", "-o", str(exe),
                        str(td / "g.cpp")], check=True)
        rng = run(exe, "rng").split()
        shipped = REF / "nbodyConfig.txt"
        cfg_out = run(exe, "config", str(shipped))
        (td / "quirky.txt").write_text(QUIRKY)
        quirky_out = run(exe, "config", str(td / "quirky.txt"))
        init_out = run(exe, "init", str(shipped))
        vec = run(exe, "vec2f").split("\n")

    def cfg(text):
        echo, vals = text.split("@@ ")
        v = vals.strip()
        path = v[v.index("[") + 1:v.rindex("]")]
        nums = v[:v.index("[")].split()
        keys = ["particleCount", "totalIterations", "save_Image_Every_Xth_Iteration", "timestep",
                "minRandBodyMass", "maxRandBodyMass", "minRadius", "maxRadius", "growthRate", "imgWidth",
                "imgHeight", "fieldWidth", "fieldHeight"]
        return {"echo": echo, "values": dict(zip(keys, nums)), "imagePath": path}

    init_lines = init_out.split("\n")
    head = [l for l in init_lines if l.startswith("@@ ")][0].split()
    bodies = {l.split()[1]: l.split()[2:] for l in init_lines if l.startswith("## ")}
    golden = {
        "generated_by": "tools/make_golden_host.py (reference headers compiled in the build container)",
        "rng": {"seed1024_ival64": rng[0:8], "seed1024_fval": rng[8:16], "seed7_ival64": rng[16:20]},
        "config_shipped": cfg(cfg_out),
        "config_quirky": {"text": QUIRKY, **cfg(quirky_out)},
        "init_shipped": {"n": int(head[1]), "fnv1a64_block": head[2], "sum_x": head[3], "sum_y": head[4],
                         "sum_m": head[5], "sum_r": head[6], "bodies_xymr_bits": bodies},
        "vec2f": [l for l in vec if l],
    }
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_text(json.dumps(golden, indent=1) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
