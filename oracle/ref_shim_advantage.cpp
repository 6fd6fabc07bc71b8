/* oracle/_ref shim for compute_puff_advantage -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Compiles the UNMODIFIED reference source where it lies under /root/reference
 * (pufferlib/extensions/pufferlib.cpp: puff_advantage_row :28-41, puff_advantage :63-72, the CPU twin of
 * extensions/cuda/pufferlib.cu) with the reference's own C++ flags (setup.py:110-112: -O3, no -march, no
 * fast-math) against the torch headers of this image, and exposes its row loop through a plain C ABI.
 * Nothing of the reference is copied here: this file only includes and calls it. */
#include "pufferlib.cpp"

extern "C" void ref_puff_advantage(float *values, float *rewards, float *dones, float *importance, float *advantages,
                                   float gamma, float lambda, float rho_clip, float c_clip, int num_steps, int horizon) {
    pufferlib::puff_advantage(values, rewards, dones, importance, advantages, gamma, lambda, rho_clip, c_clip, num_steps, horizon);
}
