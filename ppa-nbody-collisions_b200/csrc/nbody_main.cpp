// nbody_main.cpp -- drop-in replacement of the reference executable (main(), src/nbody.cu:373-551)
// on top of the C ABI in include/nbody_b200.h.
//
// With no arguments it behaves like the reference: reads ./nbodyConfig.txt, prints the same banner and
// echo lines, generates the seed-1024 initial conditions, runs totalIterations steps, writes
// <imagePath>/iteration_<k>.ppm (binary P5) on the reference's schedule and prints "Time taken".
// The reference ignores argv (:381-383), so every option below is additive:
//   --config PATH        config file (default nbodyConfig.txt)
//   --coverage MODE      reference (default: the exact pair set of src/nbody.cu:182-207) | full (all pairs)
//   --steps N            override totalIterations
//   --seed S             override the seed (default 1024, :403)
//   --scenario KIND      square (default, :401-416) | disc | two-galaxy
//   --extent R           disc radius for disc / two-galaxy (default: fieldWidth)
//   --softening EPS      opt-in Plummer softening length for the forces (not reference behaviour)
//   --merge MODE         reference (default, src/nbody.cu:215-226) | conserving (opt-in lowest-index merge that
//                        conserves mass and momentum; not reference behaviour)
//   --one-sided          never use the two-sided (pair-halving) force kernel (full coverage, >= 12288 bodies)
//   --no-images          skip rendering and image files
//   --render MODE        reference (default: the bodies the reference's stale launch grid draws, src/nbody.cu:473,535:
//                        those below 128 * floor(n_before_the_step / 128)) | all (every live body)
//   --dump-state PATH    write the final BodiesData block (int32 n, then 24 n bytes)
//   --resume PATH        start from a --dump-state file instead of generating initial conditions
//                        (checkpoint / resume: the reference has none)
//   --dump-events PATH   write the collision event list as CSV (step,i,j,kind)
//   --device D           CUDA device ordinal (of rank 0 with --gpus)
//   --gpus N             shard the step over N GPUs of this node: one host thread and one library context per GPU
//                        (devices D .. D + N - 1), NCCL between them (nb_comm_unique_id / nb_comm_init).  Rank 0 prints,
//                        draws and dumps; the event list is merged from all ranks.  Results do not depend on N for the
//                        two-sided kernel (all-pairs coverage from 12288 bodies on): its force sums are exact integers
#include <sys/time.h>

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "nbody_b200.h"

static double now_s()                                   // jbutil::gettime, include/jbutil.h:98-104
{
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return (double)tv.tv_sec + (double)tv.tv_usec * 1E-6;
}

static void die(nb_ctx *ctx, const char *what, int rc)
{
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, nb_last_error(ctx));
    exit(1);
}

int main(int argc, char **argv)
{
    const double start = now_s();
    std::string config_path = "nbodyConfig.txt", dump_state, dump_events, resume, scenario = "square";
    int coverage = NB_COVERAGE_REFERENCE, steps_override = -1, device = 0, gpus = 1;
    unsigned long long seed = 1024;
    double extent = 0, softening = 0;
    bool images = true, conserving = false, one_sided = false, render_all = false;
    for (int a = 1; a < argc; ++a) {
        const std::string opt = argv[a];
        auto need = [&](const char *name) -> const char * {
            if (a + 1 >= argc) {
                fprintf(stderr, "%s needs a value\n", name);
                exit(2);
            }
            return argv[++a];
        };
        if (opt == "--config") config_path = need("--config");
        else if (opt == "--coverage") {
            const std::string v = need("--coverage");
            if (v == "reference") coverage = NB_COVERAGE_REFERENCE;
            else if (v == "full") coverage = NB_COVERAGE_FULL;
            else { fprintf(stderr, "unknown coverage %s\n", v.c_str()); return 2; }
        }
        else if (opt == "--steps") steps_override = atoi(need("--steps"));
        else if (opt == "--seed") seed = strtoull(need("--seed"), nullptr, 10);
        else if (opt == "--scenario") scenario = need("--scenario");
        else if (opt == "--extent") extent = atof(need("--extent"));
        else if (opt == "--softening") softening = atof(need("--softening"));
        else if (opt == "--merge") conserving = std::string(need("--merge")) == "conserving";
        else if (opt == "--one-sided") one_sided = true;
        else if (opt == "--no-images") images = false;
        else if (opt == "--render") render_all = std::string(need("--render")) == "all";
        else if (opt == "--dump-state") dump_state = need("--dump-state");
        else if (opt == "--dump-events") dump_events = need("--dump-events");
        else if (opt == "--resume") resume = need("--resume");
        else if (opt == "--device") device = atoi(need("--device"));
        else if (opt == "--gpus") gpus = atoi(need("--gpus"));
        else { fprintf(stderr, "unknown option %s\n", opt.c_str()); return 2; }
    }

    fputs("Running simulation with the following settings:\n", stdout);          // :376
    fflush(stdout);
    nb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    if (nb_config_parse(config_path.c_str(), &cfg, 1) != NB_OK) return 1;        // :377 (exit(1) paths)
    fputs("=====================\n", stdout);                                     // :378
    int n0 = cfg.particleCount;
    std::vector<float> resumed;
    if (!resume.empty()) {
        FILE *f = fopen(resume.c_str(), "rb");
        int32_t n32 = 0;
        if (!f || fread(&n32, sizeof(n32), 1, f) != 1 || n32 <= 0) { fprintf(stderr, "cannot read %s\n", resume.c_str()); return 1; }
        resumed.resize((size_t)6 * n32);
        if (fread(resumed.data(), 24, (size_t)n32, f) != (size_t)n32) { fprintf(stderr, "%s is truncated\n", resume.c_str()); return 1; }
        fclose(f);
        n0 = n32;
    }
    const int total = steps_override >= 0 ? steps_override : cfg.totalIterations;
    const int every = cfg.save_Image_Every_Xth_Iteration;
    if (n0 <= 0) {
        fprintf(stderr, "particleCount must be positive\n");
        return 1;
    }
    printf("Bodies: %d\n", n0);                                                    // :399
    fflush(stdout);

    std::vector<float> block((size_t)6 * n0);
    nb_scenario sc;
    memset(&sc, 0, sizeof(sc));
    sc.kind = scenario == "disc" ? NB_SCENARIO_DISC : scenario == "two-galaxy" ? NB_SCENARIO_TWO_GALAXY : NB_SCENARIO_SQUARE;
    sc.n = n0;
    sc.seed = seed;
    sc.field_w = cfg.fieldWidth;
    sc.field_h = cfg.fieldHeight;
    sc.min_mass = cfg.minRandBodyMass;
    sc.max_mass = cfg.maxRandBodyMass;
    sc.min_radius = cfg.minRadius;
    sc.max_radius = cfg.maxRadius;
    sc.extent = extent > 0 ? extent : (double)cfg.fieldWidth;
    if (!resumed.empty()) {
        block = resumed;
    } else if (nb_generate(&sc, block.data()) != NB_OK) {
        fprintf(stderr, "invalid scenario\n");
        return 1;
    }

    if (gpus < 1) gpus = 1;
    unsigned char comm_id[NB_UNIQUE_ID_BYTES];
    std::atomic<int> comm_ready{0};
    int saved_stdout = -1;
    if (gpus > 1) {
        // NCCL may print its version banner on stdout while the communicator is set up; stdout is the reference's
        // output format, so it points at stderr until every rank has joined
        fflush(stdout);
        saved_stdout = dup(1);
        dup2(2, 1);
        const int rc = nb_comm_unique_id(comm_id);
        if (rc != NB_OK) die(nullptr, "nb_comm_unique_id", rc);
    }
    const bool do_images = images && every > 0 && cfg.imgWidth > 0 && cfg.imgHeight > 0;
    std::vector<std::vector<nb_event>> events(gpus);       // per rank: the rows it finished

    // One rank = one GPU = one library context, driven by its own host thread.  Every rank makes the same calls (the
    // collectives inside nb_step keep them in step); rank 0 alone talks to the outside world.
    auto run_rank = [&](const int rank) {
        nb_params par;
        memset(&par, 0, sizeof(par));
        par.n_max = n0;
        par.dt = cfg.timestep;
        par.growth = cfg.growthRate;
        par.field_w = cfg.fieldWidth;
        par.field_h = cfg.fieldHeight;
        par.coverage = coverage;
        par.device = device + rank;
        par.rank = rank;
        par.world = gpus;
        par.softening = (float)softening;
        if (conserving) par.flags |= NB_FLAG_MERGE_CONSERVING;
        if (one_sided) par.flags |= NB_FLAG_ONE_SIDED;
        par.event_capacity = dump_events.empty() ? 0 : 1 << 22;
        nb_ctx *ctx = nullptr;
        int rc = nb_create(&ctx, &par);
        if (rc != NB_OK) die(nullptr, "nb_create", rc);
        if (gpus > 1) {
            if ((rc = nb_comm_init(ctx, comm_id)) != NB_OK) die(ctx, "nb_comm_init", rc);
            ++comm_ready;
            if (rank == 0) {
                while (comm_ready.load() < gpus) std::this_thread::yield();
                fflush(stdout);
                dup2(saved_stdout, 1);
                close(saved_stdout);
            }
        }
        if ((rc = nb_upload(ctx, block.data(), n0)) != NB_OK) die(ctx, "nb_upload", rc);

        const bool lead = rank == 0;
        const bool want_events = !dump_events.empty();
        std::vector<nb_event> evbuf(want_events ? (1 << 22) : 0);
        std::vector<uint8_t> img(lead && do_images ? (size_t)cfg.imgWidth * cfg.imgHeight : 0);
        int pending = -1;                               // iteration whose image waits to be saved
        for (int it = 0; it < total; ++it) {
            int n_before = 0;                           // the reference draws with the grid of the step it has just done
            if (lead && do_images && it % every == 0 && !render_all && (rc = nb_num_bodies(ctx, &n_before)) != NB_OK)
                die(ctx, "nb_num_bodies", rc);
            if ((rc = nb_step(ctx, 1)) != NB_OK) die(ctx, "nb_step", rc);
            if (lead && do_images) {
                // the image rendered after iteration k is written during iteration k + 1 (:513-522)
                if (pending >= 0 && (it - 1) % every == 0) {
                    const std::string path = std::string(cfg.imagePath) + "/iteration_" + std::to_string(pending) + ".ppm";
                    printf("Saving (%dx%d) to disk\n", cfg.imgWidth, cfg.imgHeight);  // :356
                    fflush(stdout);
                    if (nb_write_pgm(path.c_str(), img.data(), cfg.imgWidth, cfg.imgHeight) != NB_OK) {
                        fprintf(stderr, "Error writing image to file:%s\nEnsure the the folder exists\n", path.c_str());   // :365-369
                        exit(1);
                    }
                    pending = -1;
                }
                if (it % every == 0) {                  // :529-539
                    const int grid = render_all ? 0x7fffffff : 128 * (n_before < 128 ? 1 : n_before / 128);              // :473,535
                    if ((rc = nb_render_grid(ctx, img.data(), cfg.imgWidth, cfg.imgHeight, grid)) != NB_OK) die(ctx, "nb_render", rc);
                    pending = it;
                }
            }
            if (want_events && (it % 64 == 63 || it == total - 1)) {
                int cnt = 0;
                if ((rc = nb_events(ctx, evbuf.data(), (int)evbuf.size(), &cnt)) != NB_OK) die(ctx, "nb_events", rc);
                events[rank].insert(events[rank].end(), evbuf.begin(), evbuf.begin() + cnt);
            }
        }
        if ((rc = nb_sync(ctx)) != NB_OK) die(ctx, "nb_sync", rc);                // CUDA_SYNC_CHECK, :546
        if (lead && !dump_state.empty()) {
            int n = 0;
            if ((rc = nb_download(ctx, block.data(), n0, &n)) != NB_OK) die(ctx, "nb_download", rc);
            FILE *f = fopen(dump_state.c_str(), "wb");
            if (!f) { fprintf(stderr, "cannot open %s\n", dump_state.c_str()); exit(1); }
            const int32_t n32 = n;
            fwrite(&n32, sizeof(n32), 1, f);
            fwrite(block.data(), 24, (size_t)n, f);
            fclose(f);
        }
        nb_destroy(ctx);
    };
    if (gpus == 1) {
        run_rank(0);
    } else {
        std::vector<std::thread> threads;
        for (int r = 0; r < gpus; ++r) threads.emplace_back(run_rank, r);
        for (std::thread &t : threads) t.join();
    }
    if (!dump_events.empty()) {
        // every rank's list is sorted by (step, row, visit order) and the ranks finish disjoint rows: a stable sort of
        // the concatenation by (step, row) is the single-GPU list
        std::vector<nb_event> all;
        for (const auto &e : events) all.insert(all.end(), e.begin(), e.end());
        std::stable_sort(all.begin(), all.end(), [](const nb_event &a, const nb_event &b) {
            return a.step != b.step ? a.step < b.step : a.i < b.i;
        });
        FILE *evf = fopen(dump_events.c_str(), "w");
        if (!evf) { fprintf(stderr, "cannot open %s\n", dump_events.c_str()); return 1; }
        fputs("step,i,j,kind\n", evf);
        for (const nb_event &e : all) fprintf(evf, "%d,%d,%d,%d\n", e.step, e.i, e.j, e.kind);
        fclose(evf);
    }
    printf("Time taken: %.4f\n", now_s() - start);                                // :548
    return 0;
}
