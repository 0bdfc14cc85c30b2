// TEST HARNESS: native-thread stress of the inference batcher (seedrl_batcher_*), the shape of
// the reference's grpc/python/ops_test.py:632-664 (10 clients x 100 single-row calls, batches
// of 5) repeated `iters` times with 3 slabs.  With native threads a caller is regularly
// descheduled between claim and commit while the others lap the slab ring -- the situation in
// which seedrl_batcher_claim used to report a spurious "would straddle two batches".
// Prints "stalls=<n> bad=<n> claim_errors=<n>"; exit code 0 iff all three are zero.
//   g++ -O2 -std=c++17 batcher_stress.cc -L<repo>/seed_rl_b200 -lseedrl_b200 -lpthread
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../include/seedrl_b200.h"

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 200;
  int stalls = 0, bad_total = 0, claim_errors = 0;
  for (int it = 0; it < iters; ++it) {
    seedrl_batcher* b = nullptr;
    size_t rows[1] = {4};
    if (seedrl_batcher_create(5, 3, 1, rows, 1, rows, 0, &b)) return 2;
    std::atomic<int> finished{0}, bad{0}, cerr{0};
    std::thread server([&] {
      for (;;) {
        int slab;
        if (seedrl_batcher_next_full(b, -1, &slab)) return;
        const int* x = (const int*)seedrl_batcher_input_ptr(b, slab, 0, 0);
        int* y = (int*)seedrl_batcher_output_ptr(b, slab, 0, 0);
        for (int i = 0; i < 5; ++i) y[i] = x[i] + 1;
        seedrl_batcher_publish(b, slab, 0);
      }
    });
    std::vector<std::thread> clients;
    for (int c = 0; c < 10; ++c) clients.emplace_back([&, c] {
      for (int i = 0; i < 100; ++i) {
        int slab, row, st;
        const int rc = seedrl_batcher_claim(b, 1, &slab, &row);
        if (rc) { if (rc != SEEDRL_ERR_CANCELLED) cerr++; return; }
        *(int*)seedrl_batcher_input_ptr(b, slab, 0, row) = c * 1000 + i;
        seedrl_batcher_commit(b, slab, 1);
        if (seedrl_batcher_wait_outputs(b, slab, &st)) { seedrl_batcher_release(b, slab); return; }
        if (*(const int*)seedrl_batcher_output_ptr(b, slab, 0, row) != c * 1000 + i + 1 || st != 0) bad++;
        seedrl_batcher_release(b, slab);
      }
      finished++;
    });
    // like the reference test: shut down once more than half the clients completed (the last
    // partially filled batch can never fill up)
    const auto t0 = std::chrono::steady_clock::now();
    while (finished.load() <= 5) {
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5)) { ++stalls; break; }
    }
    seedrl_batcher_shutdown(b);
    for (auto& t : clients) t.join();
    server.join();
    bad_total += bad.load();
    claim_errors += cerr.load();
    seedrl_batcher_destroy(b);
  }
  printf("stalls=%d bad=%d claim_errors=%d\n", stalls, bad_total, claim_errors);
  return (stalls || bad_total || claim_errors) ? 1 : 0;
}
