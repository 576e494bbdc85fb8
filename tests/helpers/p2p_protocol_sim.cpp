// CPU model of the exchange protocol of csrc/tmc_p2p.cuh (test infrastructure): N threads play the ranks, std::atomic release /
// acquire plays st.release.sys / ld.acquire.sys, and each "rank" follows exactly the kernel pair's order
//   publish(e): own buf[e & 1] <- data(rank, e);  for every peer q: flags_of_q[rank] <- e (release)
//   reduce(e):  wait own flags[q] >= e for every q (acquire);  sum_q buf_q[e & 1]  must equal  sum_q data(q, e)
// with random stalls between the steps, so that fast ranks run up to the protocol's limit ahead of slow ones.  What this checks
// is the double-buffering argument (a buffer is only rewritten two exchanges later, when every peer is provably done reading it);
// `buffers` = 1 shows the check has teeth: with a single buffer the same schedule corrupts sums.
// usage: p2p_protocol_sim <ranks> <epochs> <n> <buffers> <seed>   -> prints "mismatches <k>"
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

struct Window { std::vector<std::atomic<uint64_t>> flags; std::vector<std::vector<std::atomic<uint64_t>>> buf; };

static uint64_t data(int rank, uint64_t e, int i) { return (uint64_t)(rank + 1) * 1000003ull + e * 7919ull + (uint64_t)i; }

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 4, E = argc > 2 ? atoi(argv[2]) : 200, n = argc > 3 ? atoi(argv[3]) : 64;
  const int B = argc > 4 ? atoi(argv[4]) : 2; const unsigned seed = argc > 5 ? (unsigned)atoi(argv[5]) : 1u;
  std::vector<Window> win(N);
  for (auto& w : win) {
    w.flags = std::vector<std::atomic<uint64_t>>(N);
    for (auto& f : w.flags) f.store(0);
    w.buf.resize(B);
    for (auto& b : w.buf) { b = std::vector<std::atomic<uint64_t>>(n); for (auto& x : b) x.store(0); }
  }
  std::atomic<long> mismatches{0};
  auto rank_main = [&](int r) {
    std::mt19937 g(seed * 977u + (unsigned)r);
    auto stall = [&] { if ((g() & 7u) == 0) std::this_thread::sleep_for(std::chrono::microseconds(g() % 200)); else if (g() & 1u) std::this_thread::yield(); };
    for (uint64_t e = 1; e <= (uint64_t)E; ++e) {
      const int which = (int)(e % (uint64_t)B);
      stall();
      for (int i = 0; i < n; ++i) win[r].buf[which][i].store(data(r, e, i), std::memory_order_relaxed);      // k_p2p_publish: copy
      for (int q = 0; q < N; ++q) if (q != r) win[q].flags[r].store(e, std::memory_order_release);              // ... and signal
      stall();
      for (int q = 0; q < N; ++q) if (q != r) while (win[r].flags[q].load(std::memory_order_acquire) < e) std::this_thread::yield();   // k_p2p_reduce: wait
      stall();
      for (int i = 0; i < n; ++i) {                                                                              // ... and pull
        uint64_t s = 0, want = 0;
        for (int q = 0; q < N; ++q) { s += win[q].buf[which][i].load(std::memory_order_relaxed); want += data(q, e, i); }
        if (s != want) { mismatches++; break; }
      }
    }
  };
  std::vector<std::thread> th;
  for (int r = 0; r < N; ++r) th.emplace_back(rank_main, r);
  for (auto& t : th) t.join();
  printf("mismatches %ld\n", mismatches.load());
  return 0;
}
