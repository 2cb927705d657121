// host-side ceiling of a sparse e2e path: how fast can T threads zero-fill and scatter into a dense row buffer?
// g++ -O3 -march=native -pthread host_fill_bw.cpp -o host_fill_bw && ./host_fill_bw [MB] [threads]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? atoi(argv[1]) : 174;
  const int T = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
  const size_t n = mb * 1000000 / 4;
  float* buf = (float*)aligned_alloc(4096, n * 4);
  memset(buf, 1, n * 4);
  for (int rep = 0; rep < 6; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([=] {
        const size_t a = n * t / T, b = n * (t + 1) / T;
        memset(buf + a, 0, (b - a) * 4);
        for (size_t i = a; i < b; i += 11) buf[i] = 1.5f;  // ~9 % non-zeros
      });
    for (auto& x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %d: %zu MB zero-fill + scatter in %.3f ms = %.1f GB/s\n", T, mb, dt * 1e3, n * 4 / dt / 1e9);
  }
  return 0;
}
