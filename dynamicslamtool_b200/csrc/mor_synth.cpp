// mor_synth.cpp — seeded synthetic LiDAR sequences for MOR (C1..C5).
//
// Host-only utility (no CUDA): a beam-model ray caster over a procedural scene (ground planes,
// oriented boxes, vertical cylinders, spheres; some boxes move) with exact, consistent odometry.
// Frames are a pure function of (scenario, seed, frame index) - counter-based RNG - so the
// oracle, the GPU path, the C++ harness and bench.py all see identical bytes.
// Not part of the reference: KITTI / bag files are not available offline (BASELINE.json).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "mor_synth.h"

namespace {

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline double u01(uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
inline double hash_u01(uint64_t seed, uint64_t a, uint64_t b) { return u01(splitmix64(splitmix64(seed ^ (a * 0xD1342543DE82EF95ull)) + b)); }
inline double hash_gauss(uint64_t seed, uint64_t a, uint64_t b) {
    double u1 = hash_u01(seed, a, 2 * b), u2 = hash_u01(seed, a, 2 * b + 1);
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
}

struct Box {       // oriented (yaw) box; may move with constant world velocity, wrapped around the ego
    double c[3], h[3], yaw;
    double v[2];   // world velocity (m/s)
    double wrap;   // >0: x-position is wrapped into [ego_x - wrap, ego_x + wrap]
    double bounce; // >0: triangle-wave motion of this amplitude (m) along v direction
    float intensity;
};
struct Cyl { double c[2], r, z0, z1; float intensity; };
struct Sph { double c[3], r; float intensity; double fuzz; };
struct Plane {     // z = z0 + ax*(x-x0) + ay*(y-y0) on the rectangle [xa,xb]x[ya,yb]
    double z0, x0, y0, ax, ay, xa, xb, ya, yb; float intensity;
};

struct Pose6 { double x, y, z, yaw, pitch, roll; };

}  // namespace

struct mor_synth {
    int scenario;
    uint64_t seed;
    std::vector<double> elev;  // beam elevations (rad)
    int n_az;
    double range_min, range_max, sigma, rate_hz;
    uint32_t nominal_frames;
    std::vector<Box> boxes;
    std::vector<Cyl> cyls;
    std::vector<Sph> sphs;
    std::vector<Plane> planes;
    std::vector<Pose6> traj;  // precomputed per frame

    double terrain_z(double x, double y) const {
        double best = -1e30;
        for (const auto& p : planes)
            if (x >= p.xa && x <= p.xb && y >= p.ya && y <= p.yb) best = std::max(best, p.z0 + p.ax * (x - p.x0) + p.ay * (y - p.y0));
        return best > -1e29 ? best : 0.0;
    }
};

namespace {

void rot_zyx(const Pose6& p, double R[3][3]) {
    double cy = std::cos(p.yaw), sy = std::sin(p.yaw), cp = std::cos(p.pitch), sp = std::sin(p.pitch), cr = std::cos(p.roll), sr = std::sin(p.roll);
    R[0][0] = cy * cp; R[0][1] = cy * sp * sr - sy * cr; R[0][2] = cy * sp * cr + sy * sr;
    R[1][0] = sy * cp; R[1][1] = sy * sp * sr + cy * cr; R[1][2] = sy * sp * cr - cy * sr;
    R[2][0] = -sp;     R[2][1] = cp * sr;                R[2][2] = cp * cr;
}
void quat_zyx(const Pose6& p, double q[4]) {  // x,y,z,w
    double cy = std::cos(p.yaw * 0.5), sy = std::sin(p.yaw * 0.5), cp = std::cos(p.pitch * 0.5), sp = std::sin(p.pitch * 0.5), cr = std::cos(p.roll * 0.5), sr = std::sin(p.roll * 0.5);
    q[3] = cr * cp * cy + sr * sp * sy;
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
}

void build_scene(mor_synth& s) {
    const uint64_t sd = s.seed;
    auto rnd = [&](uint64_t a, uint64_t b) { return hash_u01(sd, 1000 + a, b); };
    const uint32_t max_frames = 4096;
    s.traj.resize(max_frames);
    if (s.scenario == 1) {
        // C1: VLP-16, indoor-scale scene for the default MOR_config.txt (trim +-3 m, gp_limit -0.5)
        for (int b = 0; b < 16; b++) s.elev.push_back((-15.0 + 2.0 * b) * M_PI / 180.0);
        s.n_az = 1800; s.range_min = 0.3; s.range_max = 100.0; s.sigma = 0.01; s.rate_hz = 10.0; s.nominal_frames = 100;
        s.planes.push_back(Plane{0, 0, 0, 0, 0, -50, 50, -50, 50, 0.10f});
        const double W = 5.2;  // walls
        s.boxes.push_back(Box{{W, 0, 1.5}, {0.1, W, 1.5}, 0, {0, 0}, 0, 0, 0.30f});
        s.boxes.push_back(Box{{-W, 0, 1.5}, {0.1, W, 1.5}, 0, {0, 0}, 0, 0, 0.30f});
        s.boxes.push_back(Box{{0, W, 1.5}, {W, 0.1, 1.5}, 0, {0, 0}, 0, 0, 0.30f});
        s.boxes.push_back(Box{{0, -W, 1.5}, {W, 0.1, 1.5}, 0, {0, 0}, 0, 0, 0.30f});
        // static furniture on both sides of the sensor's circular path (radius 3 m)
        for (int i = 0; i < 14; i++) {
            double ang = i * (2 * M_PI / 14) + 0.2 * rnd(1, i);
            double rad = (i % 2) ? 1.3 + 0.3 * rnd(2, i) : 4.4 + 0.3 * rnd(2, i);
            double hx = 0.2 + 0.2 * rnd(3, i), hy = 0.2 + 0.2 * rnd(4, i), hz = 0.4 + 0.4 * rnd(5, i);
            s.boxes.push_back(Box{{rad * std::cos(ang), rad * std::sin(ang), hz}, {hx, hy, hz}, 3.0 * rnd(6, i), {0, 0}, 0, 0, 0.5f});
        }
        for (int i = 0; i < 6; i++) {
            double ang = i * (2 * M_PI / 6) + 0.5, rad = (i % 2) ? 1.8 : 4.0;
            s.cyls.push_back(Cyl{{rad * std::cos(ang), rad * std::sin(ang)}, 0.15, 0.0, 2.5, 0.7f});
        }
        // 2 moving boxes 0.5 x 0.5 x 1.0 m, 0.5-1.0 m/s, bouncing on segments near the path
        s.boxes.push_back(Box{{2.2, -0.5, 0.5}, {0.25, 0.25, 0.5}, 0.3, {0.10, 0.75}, 0, 2.5, 0.9f});
        s.boxes.push_back(Box{{3.4, 1.6, 0.5}, {0.25, 0.25, 0.5}, 1.1, {-0.55, 0.30}, 0, 2.0, 0.9f});
        for (uint32_t f = 0; f < max_frames; f++) {
            double t = f / s.rate_hz, th = 0.1 * t;  // 0.3 m/s on a 3 m circle
            s.traj[f] = Pose6{3.0 * std::cos(th), 3.0 * std::sin(th), 0.6, th + M_PI / 2, 0.01 * std::sin(0.7 * t), 0.008 * std::sin(0.9 * t + 1.0)};
        }
    } else if (s.scenario == 2 || s.scenario == 4) {
        // C2: HDL-64E street scene; C4: same sensor over sloped / multi-plane terrain
        for (int b = 0; b < 64; b++) s.elev.push_back((2.0 - b * (26.8 / 63.0)) * M_PI / 180.0);
        s.n_az = 2083; s.range_min = 0.9; s.range_max = 80.0; s.sigma = 0.02; s.rate_hz = 10.0; s.nominal_frames = s.scenario == 2 ? 500 : 200;
        if (s.scenario == 2) {
            s.planes.push_back(Plane{0, 0, 0, 0, 0, -1e4, 1e4, -1e4, 1e4, 0.10f});
        } else {
            const double a1 = std::tan(4.0 * M_PI / 180.0), a2 = std::tan(11.0 * M_PI / 180.0);
            // two road planes meeting at a ridge every 120 m (up 4 deg for 60 m, down 11/..), plus a raised kerb plane
            for (int k = -2; k < 12; k++) {
                double xa = k * 120.0, zbase = 0.0;
                s.planes.push_back(Plane{zbase, xa, 0, a1, 0.0, xa, xa + 80.0, -1e4, 1e4, 0.10f});
                s.planes.push_back(Plane{zbase + a1 * 80.0, xa + 80.0, 0, -a1 * 2.0, 0.0, xa + 80.0, xa + 120.0, -1e4, 1e4, 0.12f});
            }
            (void)a2;
            // kerb: 0.18 m above the road on the left side
            for (int k = -2; k < 12; k++) {
                double xa = k * 120.0;
                s.planes.push_back(Plane{0.18, xa, 0, a1, 0.0, xa, xa + 80.0, 5.5, 9.5, 0.2f});
                s.planes.push_back(Plane{0.18 + a1 * 80.0, xa + 80.0, 0, -a1 * 2.0, 0.0, xa + 80.0, xa + 120.0, 5.5, 9.5, 0.2f});
            }
        }
        // street furniture generated along x in [-100, 700]
        int id = 0;
        for (double x = -100; x < 700; x += 14.0, id++) {
            for (int side = -1; side <= 1; side += 2) {
                double bx = x + 3.0 * rnd(10, id * 2 + (side > 0));
                double hx = 4.0 + 2.5 * rnd(11, id * 2 + (side > 0)), hy = 4.0 + 2.0 * rnd(12, id * 2 + (side > 0)), hz = 3.0 + 3.0 * rnd(13, id * 2 + (side > 0));
                double by = side * (13.0 + hy);
                double gz = s.terrain_z(bx, by);
                s.boxes.push_back(Box{{bx, by, gz + hz}, {hx, hy, hz}, 0.05 * (rnd(14, id) - 0.5), {0, 0}, 0, 0, 0.35f});
            }
        }
        id = 0;
        for (double x = -100; x < 700; x += 17.0, id++) {
            for (int side = -1; side <= 1; side += 2) {
                double py = side * (6.6 + 0.4 * rnd(20, id));
                double gz = s.terrain_z(x, py);
                s.cyls.push_back(Cyl{{x + 2.0 * rnd(21, id * 2 + (side > 0)), py}, 0.12 + 0.08 * rnd(22, id), gz, gz + 4.0 + 2.0 * rnd(23, id), 0.8f});
            }
        }
        id = 0;
        for (double x = -100; x < 700; x += 11.0, id++) {  // parked cars
            if (rnd(30, id) < 0.55) continue;
            int side = rnd(31, id) < 0.5 ? -1 : 1;
            double py = side * 4.6;
            if (s.scenario == 4 && side > 0) py = 4.4;
            double gz = s.terrain_z(x, py);
            s.boxes.push_back(Box{{x, py, gz + 0.75}, {2.0 + 0.3 * rnd(32, id), 0.9, 0.75}, 0.04 * (rnd(33, id) - 0.5), {0, 0}, 0, 0, 0.6f});
        }
        id = 0;
        for (double x = -100; x < 700; x += 23.0, id++) {  // bushes / pedestrians-sized blobs
            int side = rnd(40, id) < 0.5 ? -1 : 1;
            double py = side * (8.0 + 2.0 * rnd(41, id));
            double gz = s.terrain_z(x, py);
            s.sphs.push_back(Sph{{x + 5.0 * rnd(42, id), py, gz + 0.7}, 0.7 + 0.4 * rnd(43, id), 0.45f, 0.05});
        }
        // 5 moving boxes, car-sized 4 x 1.8 x 1.5 m, 3-10 m/s (wrapped to stay within +-45 m of the ego)
        const double vs[5] = {10.0, 6.5, 9.0, -5.0, 3.0};
        const double ys[5] = {-1.9, -1.9, 1.9, 1.9, -1.9};
        const double x0[5] = {12.0, 25.0, -20.0, 40.0, -8.0};
        for (int m = 0; m < 5; m++) {
            double y = ys[m] + (m == 4 ? -2.7 : 0.0);
            s.boxes.push_back(Box{{x0[m], y, 0.75}, {2.0, 0.9, 0.75}, 0.0, {vs[m], 0.0}, 45.0, 0, 0.9f});
        }
        // ego speed 2.5 m/s = 0.25 m/frame: inside the tracker's design envelope (catch_up_distance 0.3 m is
        // compared against centroids that the reference never ego-compensates, cpp:462, cpp:636)
        const double kEgo = 2.5;
        double x = 0, y = 0;
        const double dt = 1.0 / s.rate_hz;
        for (uint32_t f = 0; f < max_frames; f++) {
            double t = f * dt;
            double yaw = 0.04 * std::sin(0.25 * t);
            double gz = s.terrain_z(x, y);
            double gz2 = s.terrain_z(x + 1.0, y);
            double pitch = s.scenario == 4 ? -std::atan(gz2 - gz) : 0.003 * std::sin(1.3 * t);
            s.traj[f] = Pose6{x, y, gz + 1.73, yaw, pitch, 0.004 * std::sin(0.8 * t + 0.3)};
            const int sub = 10;  // integrate the ego speed along the heading
            for (int k = 0; k < sub; k++) {
                double tt = t + k * dt / sub, yy = 0.04 * std::sin(0.25 * tt);
                x += kEgo * std::cos(yy) * dt / sub; y += kEgo * std::sin(yy) * dt / sub;
            }
        }
    } else {
        // C3: 128-beam, dense clutter inside a walled yard
        for (int b = 0; b < 128; b++) s.elev.push_back((22.5 - b * (45.0 / 127.0)) * M_PI / 180.0);
        s.n_az = 2048; s.range_min = 0.5; s.range_max = 120.0; s.sigma = 0.015; s.rate_hz = 10.0; s.nominal_frames = 200;
        s.planes.push_back(Plane{0, 0, 0, 0, 0, -1e4, 1e4, -1e4, 1e4, 0.10f});
        const double W = 45.0;
        s.boxes.push_back(Box{{W, 0, 20}, {0.5, W, 20}, 0, {0, 0}, 0, 0, 0.30f});
        s.boxes.push_back(Box{{-W, 0, 20}, {0.5, W, 20}, 0, {0, 0}, 0, 0, 0.30f});
        s.boxes.push_back(Box{{0, W, 20}, {W, 0.5, 20}, 0, {0, 0}, 0, 0, 0.30f});
        s.boxes.push_back(Box{{0, -W, 20}, {W, 0.5, 20}, 0, {0, 0}, 0, 0, 0.30f});
        for (int i = 0; i < 200; i++) {
            double ang = 2 * M_PI * rnd(50, i), rad = 4.0 + 26.0 * std::sqrt(rnd(51, i));
            double cx = rad * std::cos(ang), cy = rad * std::sin(ang);
            if (i % 3 == 0) s.cyls.push_back(Cyl{{cx, cy}, 0.15 + 0.5 * rnd(52, i), 0.0, 1.0 + 5.0 * rnd(53, i), 0.7f});
            else if (i % 3 == 1) {
                double hz = 0.3 + 1.5 * rnd(54, i);
                s.boxes.push_back(Box{{cx, cy, hz}, {0.3 + 1.2 * rnd(55, i), 0.3 + 1.2 * rnd(56, i), hz}, 3.0 * rnd(57, i), {0, 0}, 0, 0, 0.5f});
            } else s.sphs.push_back(Sph{{cx, cy, 0.5 + 1.5 * rnd(58, i)}, 0.5 + 1.0 * rnd(59, i), 0.45f, 0.08});
        }
        for (int m = 0; m < 6; m++) {
            double ang = m * 1.05;
            s.boxes.push_back(Box{{9.0 * std::cos(ang), 9.0 * std::sin(ang), 0.8}, {0.9, 0.5, 0.8}, ang, {1.5 * std::cos(ang + 1.57), 1.5 * std::sin(ang + 1.57)}, 0, 6.0, 0.9f});
        }
        for (uint32_t f = 0; f < max_frames; f++) {
            double t = f / s.rate_hz, th = 0.2 * t;  // 2 m/s on a 10 m circle... kept small: 2 m radius, 0.4 m/s
            s.traj[f] = Pose6{2.0 * std::cos(th), 2.0 * std::sin(th), 1.8, th + M_PI / 2, 0.004 * std::sin(0.9 * t), 0.004 * std::sin(1.1 * t)};
        }
    }
}

struct FrameCtx {
    const mor_synth* s;
    Pose6 pose; double R[3][3];
    double t;
    std::vector<Box> boxes;  // positions resolved at time t, culled
    std::vector<Cyl> cyls;
    std::vector<Sph> sphs;
};

inline bool ray_sphere_reject(const double o[3], const double d[3], const double c[3], double r) {
    double oc[3] = {c[0] - o[0], c[1] - o[1], c[2] - o[2]};
    double b = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2];
    if (b < -r) return true;
    double perp2 = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - b * b;
    return perp2 > r * r;
}

void cast_range(const FrameCtx& fc, uint32_t frame, int ray_lo, int ray_hi, std::vector<float>& out) {
    const mor_synth& s = *fc.s;
    const int nb = (int)s.elev.size();
    const double o[3] = {fc.pose.x, fc.pose.y, fc.pose.z};
    for (int ray = ray_lo; ray < ray_hi; ray++) {
        const int az_i = ray / nb, b = ray % nb;
        const double az = -(az_i + 0.5) * (2 * M_PI / s.n_az);  // Velodyne spins clockwise
        const double ce = std::cos(s.elev[b]);
        const double ds[3] = {ce * std::cos(az), ce * std::sin(az), std::sin(s.elev[b])};
        double d[3];
        for (int i = 0; i < 3; i++) d[i] = fc.R[i][0] * ds[0] + fc.R[i][1] * ds[1] + fc.R[i][2] * ds[2];
        double best = s.range_max; float inten = 0.f; double fuzz = 0.0;
        for (const auto& p : s.planes) {
            // z = z0 + ax (x-x0) + ay (y-y0)  <=>  n.p = k with n = (-ax,-ay,1)
            double denom = d[2] - p.ax * d[0] - p.ay * d[1];
            if (denom > -1e-9) continue;  // only hit from above
            double num = (p.z0 - p.ax * p.x0 - p.ay * p.y0) - (o[2] - p.ax * o[0] - p.ay * o[1]);
            double tt = num / denom;
            if (tt <= s.range_min || tt >= best) continue;
            double hx = o[0] + tt * d[0], hy = o[1] + tt * d[1];
            if (hx < p.xa || hx > p.xb || hy < p.ya || hy > p.yb) continue;
            best = tt; inten = p.intensity; fuzz = 0;
        }
        for (const auto& bx : fc.boxes) {
            double rad = std::sqrt(bx.h[0] * bx.h[0] + bx.h[1] * bx.h[1] + bx.h[2] * bx.h[2]);
            if (ray_sphere_reject(o, d, bx.c, rad)) continue;
            double cy = std::cos(bx.yaw), sy = std::sin(bx.yaw);
            double rel[3] = {o[0] - bx.c[0], o[1] - bx.c[1], o[2] - bx.c[2]};
            double lo3[3] = {cy * rel[0] + sy * rel[1], -sy * rel[0] + cy * rel[1], rel[2]};
            double ld[3] = {cy * d[0] + sy * d[1], -sy * d[0] + cy * d[1], d[2]};
            double t0 = 0, t1 = best; bool hit = true;
            for (int a = 0; a < 3 && hit; a++) {
                if (std::fabs(ld[a]) < 1e-12) { if (std::fabs(lo3[a]) > bx.h[a]) hit = false; continue; }
                double ta = (-bx.h[a] - lo3[a]) / ld[a], tb = (bx.h[a] - lo3[a]) / ld[a];
                if (ta > tb) std::swap(ta, tb);
                t0 = std::max(t0, ta); t1 = std::min(t1, tb);
                if (t0 > t1) hit = false;
            }
            if (hit && t0 > s.range_min && t0 < best) { best = t0; inten = bx.intensity; fuzz = 0; }
        }
        for (const auto& c : fc.cyls) {
            double cc[3] = {c.c[0], c.c[1], 0.5 * (c.z0 + c.z1)};
            double rad = std::sqrt(c.r * c.r + 0.25 * (c.z1 - c.z0) * (c.z1 - c.z0));
            if (ray_sphere_reject(o, d, cc, rad)) continue;
            double ox = o[0] - c.c[0], oy = o[1] - c.c[1];
            double A = d[0] * d[0] + d[1] * d[1], B = ox * d[0] + oy * d[1], C = ox * ox + oy * oy - c.r * c.r;
            if (A < 1e-14) continue;
            double disc = B * B - A * C;
            if (disc < 0) continue;
            double tt = (-B - std::sqrt(disc)) / A;
            if (tt <= s.range_min || tt >= best) continue;
            double hz = o[2] + tt * d[2];
            if (hz < c.z0 || hz > c.z1) continue;
            best = tt; inten = c.intensity; fuzz = 0;
        }
        for (const auto& sp : fc.sphs) {
            if (ray_sphere_reject(o, d, sp.c, sp.r)) continue;
            double oc[3] = {o[0] - sp.c[0], o[1] - sp.c[1], o[2] - sp.c[2]};
            double B = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2], C = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - sp.r * sp.r;
            double disc = B * B - C;
            if (disc < 0) continue;
            double tt = -B - std::sqrt(disc);
            if (tt <= s.range_min || tt >= best) continue;
            best = tt; inten = sp.intensity; fuzz = sp.fuzz;
        }
        if (best >= s.range_max) continue;
        double sig = std::sqrt(s.sigma * s.sigma + fuzz * fuzz);
        double r = best + sig * hash_gauss(s.seed, ((uint64_t)frame << 24) ^ (uint64_t)ray, 7);
        float rec[4] = {(float)(ds[0] * r), (float)(ds[1] * r), (float)(ds[2] * r),
                        inten + 0.02f * (float)hash_u01(s.seed, ((uint64_t)frame << 24) ^ (uint64_t)ray, 9)};
        out.insert(out.end(), rec, rec + 4);
    }
}

}  // namespace

extern "C" {

int mor_synth_create(int scenario, uint64_t seed, mor_synth** out) {
    if (!out || scenario < 1 || scenario > 4) return 5;
    mor_synth* s = new mor_synth();
    s->scenario = scenario; s->seed = seed;
    build_scene(*s);
    *out = s;
    return 0;
}
int mor_synth_destroy(mor_synth* s) { delete s; return 0; }

int mor_synth_info(const mor_synth* s, uint32_t* max_points, uint32_t* nominal_frames, double* rate_hz) {
    if (!s) return 5;
    if (max_points) *max_points = (uint32_t)(s->elev.size() * s->n_az);
    if (nominal_frames) *nominal_frames = s->nominal_frames;
    if (rate_hz) *rate_hz = s->rate_hz;
    return 0;
}

int mor_synth_frame(const mor_synth* s, uint32_t frame, float* xyzi, uint32_t cap_points, uint32_t* n_points, double pose7[7], int n_threads) {
    if (!s || !xyzi || !n_points || !pose7 || frame >= s->traj.size()) return 5;
    FrameCtx fc;
    fc.s = s; fc.pose = s->traj[frame]; fc.t = frame / s->rate_hz;
    rot_zyx(fc.pose, fc.R);
    const double cull = s->range_max + 1.0;
    for (auto b : s->boxes) {
        if (b.bounce > 0) {  // triangle wave along v
            double sp = std::sqrt(b.v[0] * b.v[0] + b.v[1] * b.v[1]);
            double dist = sp * fc.t, period = 2 * b.bounce, ph = std::fmod(dist, period);
            double off = ph < b.bounce ? ph : period - ph;
            b.c[0] += b.v[0] / sp * off; b.c[1] += b.v[1] / sp * off;
        } else if (b.v[0] != 0 || b.v[1] != 0) {
            b.c[0] += b.v[0] * fc.t; b.c[1] += b.v[1] * fc.t;
            if (b.wrap > 0) {
                double rel = b.c[0] - fc.pose.x;
                rel = rel - 2 * b.wrap * std::floor((rel + b.wrap) / (2 * b.wrap));
                b.c[0] = fc.pose.x + rel;
            }
            b.c[2] = s->terrain_z(b.c[0], b.c[1]) + b.h[2];
        }
        double dx = b.c[0] - fc.pose.x, dy = b.c[1] - fc.pose.y;
        double rad = std::sqrt(b.h[0] * b.h[0] + b.h[1] * b.h[1]);
        if (std::sqrt(dx * dx + dy * dy) - rad > cull) continue;
        fc.boxes.push_back(b);
    }
    for (const auto& c : s->cyls) {
        double dx = c.c[0] - fc.pose.x, dy = c.c[1] - fc.pose.y;
        if (std::sqrt(dx * dx + dy * dy) - c.r > cull) continue;
        fc.cyls.push_back(c);
    }
    for (const auto& sp : s->sphs) {
        double dx = sp.c[0] - fc.pose.x, dy = sp.c[1] - fc.pose.y;
        if (std::sqrt(dx * dx + dy * dy) - sp.r > cull) continue;
        fc.sphs.push_back(sp);
    }
    const int n_rays = (int)s->elev.size() * s->n_az;
    int nt = std::max(1, std::min(n_threads, 64));
    std::vector<std::vector<float>> parts(nt);
    if (nt == 1) cast_range(fc, frame, 0, n_rays, parts[0]);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) {
            int lo = (int)((int64_t)n_rays * t / nt), hi = (int)((int64_t)n_rays * (t + 1) / nt);
            th.emplace_back([&, t, lo, hi] { cast_range(fc, frame, lo, hi, parts[t]); });
        }
        for (auto& t : th) t.join();
    }
    size_t total = 0;
    for (auto& p : parts) total += p.size() / 4;
    *n_points = (uint32_t)total;
    if (total > cap_points) return 6;
    size_t off = 0;
    for (auto& p : parts) { std::memcpy(xyzi + off, p.data(), p.size() * sizeof(float)); off += p.size(); }
    pose7[0] = fc.pose.x; pose7[1] = fc.pose.y; pose7[2] = fc.pose.z;
    quat_zyx(fc.pose, pose7 + 3);
    return 0;
}

}  // extern "C"
