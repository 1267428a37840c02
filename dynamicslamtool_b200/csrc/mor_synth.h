/* mor_synth.h — C ABI of the seeded synthetic LiDAR sequence generator (host only).
 * Scenarios: 1 = C1 VLP-16 indoor (default MOR_config.txt), 2 = C2 HDL-64E street,
 * 3 = C3 128-beam clutter, 4 = C4 HDL-64E over sloped / multi-plane terrain.
 * A frame is a pure function of (scenario, seed, frame index). Returns 0 on success. */
#ifndef MOR_SYNTH_H
#define MOR_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mor_synth mor_synth;
int mor_synth_create(int scenario, uint64_t seed, mor_synth** out);
int mor_synth_destroy(mor_synth* s);
int mor_synth_info(const mor_synth* s, uint32_t* max_points, uint32_t* nominal_frames, double* rate_hz);
/* xyzi: cap_points * 4 floats (x,y,z,intensity in the sensor frame, 16 B/point);
 * pose7: sensor pose in the world: position x,y,z + orientation quaternion x,y,z,w. */
int mor_synth_frame(const mor_synth* s, uint32_t frame, float* xyzi, uint32_t cap_points,
                    uint32_t* n_points, double pose7[7], int n_threads);
#ifdef __cplusplus
}
#endif
#endif
