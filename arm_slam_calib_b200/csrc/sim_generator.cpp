// Host-side synthetic problem generator: restates, from scratch, what the reference does to
// build its simulated calibration graph, so that factor/landmark association and indexing are
// identical.  Reference behaviour followed (not copied):
//   landmark grid ............. ArmSlamCalib::CreateSimulation      src/ArmSlamCalib.cpp:654-665
//   random-walk trajectory .... ArmSlamCalib::CreateSimTrajectory   src/ArmSlamCalib.cpp:535-602
//   Perlin term ............... ArmSlamCalib::GetPerlinNoise        src/ArmSlamCalib.cpp:2296-2306
//                               PerlinNoise::{PerlinNoise,noise}    src/PerlinNoise.cpp:37-137
//   visibility test ........... GetNewLandmarksSimulated            src/ArmSlamCalib.cpp:1773-1810
//   admission + factor order .. SimulateObservationsTriangulate     src/ArmSlamCalib.cpp:1812-1887
//                               SimulationStep                      src/ArmSlamCalib.cpp:307-327
//   random chain .............. RobotGenerator::GenerateRobot       src/RobotGenerator.cpp:22-147
//   utils::Rand ............... src/Utils.cpp:17-30 (libc rand())
// The reference draws from std::default_random_engine / std::normal_distribution / std::shuffle;
// using the same libstdc++ classes here reproduces its streams exactly.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include "../../include/asc.h"
#include "host_math.h"

namespace {

using asc::Xf;

double rand_unit() { return (double)rand() / (double)RAND_MAX; }            // Utils.cpp:17-20
double rand_range(double lo, double hi) { return rand_unit() * (hi - lo) + lo; }  // Utils.cpp:27-30

// Improved-Perlin lattice noise with a seeded permutation (PerlinNoise.cpp:37-137)
struct Perlin {
    std::vector<int> p;
    explicit Perlin(unsigned seed) {
        p.resize(256);
        std::iota(p.begin(), p.end(), 0);
        std::default_random_engine engine(seed);
        std::shuffle(p.begin(), p.end(), engine);
        p.insert(p.end(), p.begin(), p.end());
    }
    static double fade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }
    static double lerp(double t, double a, double b) { return a + t * (b - a); }
    static double grad(int hash, double x, double y, double z) {
        const int h = hash & 15;
        const double u = h < 8 ? x : y;
        const double v = h < 4 ? y : (h == 12 || h == 14 ? x : z);
        return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
    }
    double noise(double x, double y, double z) const {
        const int X = (int)std::floor(x) & 255, Y = (int)std::floor(y) & 255, Z = (int)std::floor(z) & 255;
        x -= std::floor(x); y -= std::floor(y); z -= std::floor(z);
        const double u = fade(x), v = fade(y), w = fade(z);
        const int A = p[X] + Y, AA = p[A] + Z, AB = p[A + 1] + Z;
        const int B = p[X + 1] + Y, BA = p[B] + Z, BB = p[B + 1] + Z;
        const double lo = lerp(v, lerp(u, grad(p[AA], x, y, z), grad(p[BA], x - 1, y, z)),
                               lerp(u, grad(p[AB], x, y - 1, z), grad(p[BB], x - 1, y - 1, z)));
        const double hi = lerp(v, lerp(u, grad(p[AA + 1], x, y, z - 1), grad(p[BA + 1], x - 1, y, z - 1)),
                               lerp(u, grad(p[AB + 1], x, y - 1, z - 1), grad(p[BB + 1], x - 1, y - 1, z - 1)));
        return (lerp(w, lo, hi) + 1.0) / 2.0;
    }
};

struct Chain {
    int n = 0;
    std::vector<Xf> p2j, c2j_inv;
    std::vector<double> axis;
};

Chain make_chain(int n, const double* t_p2j, const double* axis, const double* t_c2j) {
    Chain c;
    c.n = n;
    for (int i = 0; i < n; ++i) {
        c.p2j.push_back(asc::xf_from12(t_p2j + 12 * i));
        c.c2j_inv.push_back(asc::xf_inv(asc::xf_from12(t_c2j + 12 * i)));
        for (int k = 0; k < 3; ++k) c.axis.push_back(axis[3 * i + k]);
    }
    return c;
}

// DART convention: child = parent * T_p2j * R(axis,q) * T_c2j^-1
Xf chain_fk(const Chain& c, const double* q) {
    Xf T = asc::xf_identity();
    for (int i = 0; i < c.n; ++i)
        T = asc::xf_mul(asc::xf_mul(asc::xf_mul(T, c.p2j[i]), asc::xf_rot(&c.axis[3 * i], q[i])), c.c2j_inv[i]);
    return T;
}

}  // namespace

struct asc_sim_problem {
    int n_dof = 0, n_poses = 0, n_landmarks = 0, n_sim_landmarks = 0;
    std::vector<int32_t> pose_idx, lm_idx, lm_id;
    std::vector<int64_t> factor_order;
    std::vector<double> uv, enc, q_truth, lm_truth;
};

extern "C" {

void asc_sim_default_config(asc_sim_config* c) {
    c->trajectory_size = 1000;            // ArmSlamCalib.h:70
    c->num_landmarks_x = 25;              // :71
    c->num_landmarks_y = 25;              // :72
    c->landmark_size_x = 10;              // :73
    c->landmark_size_y = 10;              // :74
    c->landmark_z_min = -5;               // ArmSlamCalib.cpp:662
    c->landmark_z_max = 5;
    c->max_range = 0;
    c->encoder_noise = 0.1;               // ArmSlamCalib.h:79
    c->velocity_noise = 0.01;             // :80
    c->seed = 0;
    c->min_observations = 3;              // ArmSlamCalib.cpp:1860
    const Xf I = asc::xf_identity();
    asc::xf_to12(I, c->sim_extrinsic);    // ArmSlamCalib.cpp:193
    asc::xf_to12(I, c->extrinsic_guess);  // ArmSlamCalib.h:87
    c->segment_length = 0;
}

int asc_chain_fk(int32_t n, const double* t_p2j, const double* axis, const double* t_c2j, const double* q,
                 double* t_link) {
    if (n <= 0 || n > ASC_MAX_DOF) return ASC_ERR_INVALID;
    const Chain c = make_chain(n, t_p2j, axis, t_c2j);
    asc::xf_to12(chain_fk(c, q), t_link);
    return ASC_OK;
}

// Kinova Mico-like 6-DOF fixture.  The reference loads package://ada_description/robots/mico.urdf
// (src/sim.cpp:64), which is not vendored; this chain uses Kinova's published classic-DH link
// parameters and is labelled synthetic everywhere it is reported.
int asc_chain_mico(double* t_p2j, double* axis, double* t_c2j) {
    const double D1 = 0.2755, D2 = 0.2900, D3 = 0.1233, D4 = 0.0741, D5 = 0.0741, D6 = 0.1600, e2 = 0.0070;
    const double aa = 30.0 * M_PI / 180.0, sa = std::sin(aa), s2a = std::sin(2 * aa);
    const double d4b = D3 + sa / s2a * D4, d5b = sa / s2a * D4 + sa / s2a * D5, d6b = sa / s2a * D5 + D6;
    const double alpha[6] = {M_PI / 2, M_PI, M_PI / 2, 2 * aa, 2 * aa, M_PI};
    const double a[6] = {0, D2, 0, 0, 0, 0};
    const double d[6] = {D1, 0, -e2, -d4b, -d5b, -d6b};
    const double X[3] = {1, 0, 0};
    for (int i = 0; i < 6; ++i) {
        // link transform Tz(d) Tx(a) Rx(alpha) after the joint rotation == inverse(T_c2j)
        Xf L = asc::xf_rot(X, alpha[i]);
        L.t[0] = a[i];
        L.t[2] = d[i];
        asc::xf_to12(asc::xf_identity(), t_p2j + 12 * i);
        asc::xf_to12(asc::xf_inv(L), t_c2j + 12 * i);
        axis[3 * i] = 0; axis[3 * i + 1] = 0; axis[3 * i + 2] = 1;
    }
    return 6;
}

int asc_chain_generate(int32_t n_dof, double scale_d, double growth_d, double random_length_d, double random_angle_d,
                       double* t_p2j, double* axis, double* t_c2j) {
    if (n_dof <= 0 || n_dof > ASC_MAX_DOF) return -1;
    // the reference's arguments and running lengths are float (RobotGenerator.cpp:22-28, :88-89)
    const float scale = (float)scale_d, growth = (float)growth_d;
    const float random_length = (float)random_length_d, random_angle = (float)random_angle_d;
    const double X[3] = {1, 0, 0}, Y[3] = {0, 1, 0};
    float last_len = 0, link_scale = scale;
    for (int i = 0; i < n_dof; ++i) {
        const double parity = i % 2 == 0 ? 1.0 : -1.0;
        Xf p2j = asc::xf_identity(), c2j = asc::xf_identity();
        p2j.t[0] = last_len * 0.5 + link_scale * 0.125;
        c2j.t[0] = -last_len * 0.25;
        if (i == n_dof - 1) {
            const Xf twist = asc::xf_rot(X, 1.57);
            for (int k = 0; k < 9; ++k) c2j.R[k] = twist.R[k];
        }
        if (i > 0) {
            const double ang = parity * 1.57 + rand_range(-random_angle, random_angle);
            const Xf r = asc::xf_rot(parity < 0 ? Y : X, ang);
            for (int k = 0; k < 9; ++k) p2j.R[k] = r.R[k];
        } else {
            p2j = asc::xf_rot(Y, -1.57);   // assignment of Identity() also clears the translation
            c2j.t[0] = -scale * 0.5;
        }
        asc::xf_to12(p2j, t_p2j + 12 * i);
        asc::xf_to12(c2j, t_c2j + 12 * i);
        axis[3 * i] = 0; axis[3 * i + 1] = 0; axis[3 * i + 2] = 1;
        last_len = link_scale;
        link_scale *= growth;
        link_scale += rand_range(-random_length, random_length);
        link_scale = std::fmax(link_scale, 0.01f);
    }
    return n_dof;
}

int asc_sim_create(const asc_sim_config* cfg, int32_t n_dof, const double* t_p2j, const double* axis,
                   const double* t_c2j, double fx, double fy, double cx, double cy, asc_sim_problem** out) {
    if (!cfg || !out || n_dof <= 0 || n_dof > ASC_MAX_DOF || cfg->trajectory_size <= 0) return ASC_ERR_INVALID;
    const Chain chain = make_chain(n_dof, t_p2j, axis, t_c2j);
    const int N = n_dof, T = cfg->trajectory_size;
    const int nx = cfg->num_landmarks_x, ny = cfg->num_landmarks_y;

    // ---- landmarks: one rand() per grid node, x outer / y inner (ArmSlamCalib.cpp:656-665)
    std::vector<double> sim_lm((size_t)nx * ny * 3);
    const double dx = cfg->landmark_size_x / nx, dy = cfg->landmark_size_y / ny;
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y) {
            const double lx = x * dx - cfg->landmark_size_x * 0.5f;
            const double ly = y * dy - cfg->landmark_size_y * 0.5f;
            const double lz = rand_range(cfg->landmark_z_min, cfg->landmark_z_max);
            double* l = &sim_lm[((size_t)x * ny + y) * 3];
            l[0] = lx * 2; l[1] = ly * 2; l[2] = lz;
        }

    // ---- trajectory (ArmSlamCalib.cpp:535-602); joint limits +-3.0 as RobotGenerator.cpp:97-98
    const double lower = -3.0, upper = 3.0;
    std::vector<double> traj((size_t)T * N), encs((size_t)T * N);
    {
        Perlin perlin((unsigned)cfg->seed);
        std::default_random_engine generator(cfg->seed);
        std::normal_distribution<double> distribution(0, cfg->encoder_noise);
        std::normal_distribution<double> velocity_distribution(0, cfg->velocity_noise);
        const double freq = 1.0f, mag = 0.01f;   // float literals at the call site (:551)
        std::vector<double> q(N, 0.0), q_dot(N, (double)0.001f), config(N, 0.0), noise(N), next(N);
        for (int i = 0; i < T; ++i) {
            if (cfg->segment_length > 0 && i > 0 && i % cfg->segment_length == 0) {
                // [NEW] restart the walk from a random configuration inside 80% of the limits
                std::default_random_engine start_gen((unsigned)(cfg->seed + 7919 * (i / cfg->segment_length)));
                std::uniform_real_distribution<double> start(0.8 * lower, 0.8 * upper);
                for (int k = 0; k < N; ++k) { config[k] = q[k] = start(start_gen); q_dot[k] = (double)0.001f; }
            }
            for (int j = 0; j < N; ++j) {
                const double pos = q[j] + 0.25 * (i * 0.01);
                noise[j] = (perlin.noise((double)(j + 1) * 100, -(double)(j + 1) * 999, pos * freq) - 0.5) * mag;
            }
            for (int k = 0; k < N; ++k) noise[k] += distribution(generator) * 0.01;
            for (int k = 0; k < N; ++k) q_dot[k] += noise[k];
            for (int k = 0; k < N; ++k) next[k] = config[k] + q_dot[k];            // retract
            for (int k = 0; k < N; ++k) q[k] = config[k];                          // lags one step (:568)
            for (int k = 0; k < N; ++k) q_dot[k] = next[k] - config[k];            // localCoordinates
            config = next;
            for (int k = 0; k < N; ++k) {
                if (q[k] > upper || q[k] < lower) q_dot[k] *= -0.5;
                q[k] = std::fmax(std::fmin(q[k], upper), lower);
                encs[(size_t)i * N + k] = q[k] + distribution(generator);
                (void)velocity_distribution(generator);   // noisyVelocity draw keeps the stream aligned (:588)
                traj[(size_t)i * N + k] = q[k];
            }
        }
    }

    // ---- observations, admission and factor order (ArmSlamCalib.cpp:307-327, :1773-1887)
    const Xf sim_ext = asc::xf_from12(cfg->sim_extrinsic);
    const double r2max = cfg->max_range > 0 ? cfg->max_range * cfg->max_range : 0;
    struct Track { std::vector<int32_t> cfgs; std::vector<double> uvs; bool in_graph = false; };
    std::vector<Track> tracks((size_t)nx * ny);
    struct Obs { int32_t pose, lm; double u, v; int64_t order; };
    std::vector<Obs> obs;
    int64_t factor_counter = 1;   // factor #0 is PriorFactor<Pose3>(K0) (:685)
    std::vector<char> admitted((size_t)nx * ny, 0);
    for (int t = 0; t < T; ++t) {
        ++factor_counter;   // EncoderFactor(q_t) (:316)
        const Xf cam = asc::xf_mul(chain_fk(chain, &traj[(size_t)t * N]), sim_ext);
        int x0 = 0, x1 = nx, y0 = 0, y1 = ny;
        if (r2max > 0) {   // [NEW] grid culling; does not change the ascending-id visit order
            x0 = std::max(0, (int)std::floor(((cam.t[0] - cfg->max_range) / 2 + cfg->landmark_size_x * 0.5f) / dx));
            x1 = std::min(nx, (int)std::ceil(((cam.t[0] + cfg->max_range) / 2 + cfg->landmark_size_x * 0.5f) / dx) + 1);
            y0 = std::max(0, (int)std::floor(((cam.t[1] - cfg->max_range) / 2 + cfg->landmark_size_y * 0.5f) / dy));
            y1 = std::min(ny, (int)std::ceil(((cam.t[1] + cfg->max_range) / 2 + cfg->landmark_size_y * 0.5f) / dy) + 1);
        }
        for (int x = x0; x < x1; ++x)
            for (int y = y0; y < y1; ++y) {
                const size_t id = (size_t)x * ny + y;
                double pc[3];
                asc::xf_transform_to(cam, &sim_lm[id * 3], pc);
                if (!(pc[2] > 0)) continue;                                   // projectSafe (:1784)
                if (r2max > 0 && pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2] > r2max) continue;
                const double u = fx * (pc[0] / pc[2]) + cx, v = fy * (pc[1] / pc[2]) + cy;
                if (!(u > 0 && v > 0 && u < cx * 2 && v < cy * 2)) continue;  // :1786
                Track& tr = tracks[id];
                tr.cfgs.push_back(t);
                tr.uvs.push_back(u);
                tr.uvs.push_back(v);
                if (!tr.in_graph && (int)tr.cfgs.size() > cfg->min_observations) {
                    tr.in_graph = true;
                    admitted[id] = 1;
                    ++factor_counter;   // PriorFactor<Point3> (:1870)
                    for (size_t k = 0; k < tr.cfgs.size(); ++k)
                        obs.push_back({tr.cfgs[k], (int32_t)id, tr.uvs[2 * k], tr.uvs[2 * k + 1], factor_counter++});
                } else if (tr.in_graph) {
                    obs.push_back({t, (int32_t)id, u, v, factor_counter++});
                }
            }
    }

    // ---- compact landmark index = rank of the id among admitted landmarks (Values iterate by key)
    asc_sim_problem* s = new asc_sim_problem();
    s->n_dof = N;
    s->n_poses = T;
    s->n_sim_landmarks = nx * ny;
    std::vector<int32_t> compact((size_t)nx * ny, -1);
    for (size_t id = 0; id < admitted.size(); ++id)
        if (admitted[id]) {
            compact[id] = (int32_t)s->lm_id.size();
            s->lm_id.push_back((int32_t)id);
            for (int k = 0; k < 3; ++k) s->lm_truth.push_back(sim_lm[id * 3 + k]);
        }
    s->n_landmarks = (int)s->lm_id.size();
    s->pose_idx.reserve(obs.size()); s->lm_idx.reserve(obs.size());
    s->uv.reserve(obs.size() * 2); s->factor_order.reserve(obs.size());
    for (const Obs& o : obs) {
        s->pose_idx.push_back(o.pose);
        s->lm_idx.push_back(compact[o.lm]);
        s->uv.push_back(o.u);
        s->uv.push_back(o.v);
        s->factor_order.push_back(o.order);
    }
    s->enc = encs;
    s->q_truth = traj;
    *out = s;
    return ASC_OK;
}

void asc_sim_destroy(asc_sim_problem* s) { delete s; }

int asc_sim_sizes(const asc_sim_problem* s, int32_t* n_poses, int32_t* n_landmarks, int64_t* n_obs,
                  int32_t* n_sim_landmarks) {
    if (!s) return ASC_ERR_INVALID;
    if (n_poses) *n_poses = s->n_poses;
    if (n_landmarks) *n_landmarks = s->n_landmarks;
    if (n_obs) *n_obs = (int64_t)s->pose_idx.size();
    if (n_sim_landmarks) *n_sim_landmarks = s->n_sim_landmarks;
    return ASC_OK;
}

int asc_sim_export(const asc_sim_problem* s, int32_t* pose_idx, int32_t* lm_idx, double* uv, double* enc,
                   double* q_truth, double* lm_truth, int32_t* lm_id, int64_t* factor_order) {
    if (!s) return ASC_ERR_INVALID;
    auto cp = [](auto* dst, const auto& v) { if (dst) std::copy(v.begin(), v.end(), dst); };
    cp(pose_idx, s->pose_idx); cp(lm_idx, s->lm_idx); cp(uv, s->uv); cp(enc, s->enc);
    cp(q_truth, s->q_truth); cp(lm_truth, s->lm_truth); cp(lm_id, s->lm_id); cp(factor_order, s->factor_order);
    return ASC_OK;
}

void asc_srand(unsigned seed) { srand(seed); }   // sim_genrobot.cpp:60

}  // extern "C"
