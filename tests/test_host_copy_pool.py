"""Host logic of the host-pointer FD call's staging of pageable caller buffers (ilqg-mujoco_b200/csrc/host_copy.h: BounceCrew, CopyPool)
driven on the CPU by a small C++ harness — no GPU, no CUDA: the header is plain C++.

The harness plays the call's own role (ilqg_fd_batch_host, ilqg.cu): per chunk it waits for the chunk's input slices (`inputs_of`), checks
that they are in place before it would issue the upload, later "lands" the chunks in order and finally checks every output byte.  Cases:
thread counts 1 / 3 / 8 / 16, 1 / 5 / 32 chunks, job sizes from 1 byte to a few MB with odd lengths and misaligned pointers (the
streaming-store path needs 16-byte alignment and falls back to memcpy), chunks without jobs, a pool reused for 300 crews while its thread
count changes, and a crew destroyed before its downloads land (the early error return of the call): no worker may touch it afterwards
and nothing may hang."""
import os
import subprocess

import pytest

from conftest import ROOT

HARNESS = r"""
#include "host_copy.h"
#include <cstdio>
#include <cstdlib>
#include <random>

static std::mt19937_64 rng(12345);
static size_t rnd(size_t lo, size_t hi) { return lo + rng() % (hi - lo + 1); }

struct Buf { std::vector<char> mem; size_t off; char* p() { return mem.data() + off; } };
static Buf make(size_t bytes, bool misalign) {
    Buf b; b.mem.resize(bytes + 64); b.off = misalign ? 1 + rng() % 15 : (64 - ((uintptr_t)b.mem.data() & 63)) & 63; return b;
}

static int run_case(CopyPool& pool, int K, int nchunks, size_t maxbytes, bool abort_early) {
    std::vector<Buf> src_in, dst_in, src_out, dst_out;
    std::vector<std::vector<int>> in_idx(nchunks), out_idx(nchunks);
    BounceCrew crew;   // (declared after the buffers: destroyed first, and its destructor waits for the workers to let go of them)
    for (int c = 0; c < nchunks; c++) {
        const int nin = (int)rnd(0, 4), nout = (int)rnd(0, 2);
        for (int j = 0; j < nin + nout; j++) {
            const size_t bytes = rnd(1, maxbytes);
            Buf s = make(bytes, rng() % 3 == 0), d = make(bytes, rng() % 3 == 0);
            for (size_t i = 0; i < bytes; i++) { s.p()[i] = (char)(rng() & 0xff); d.p()[i] = 0x5a; }
            auto& S = j < nin ? src_in : src_out; auto& D = j < nin ? dst_in : dst_out;
            S.push_back(std::move(s)); D.push_back(std::move(d));
            (j < nin ? in_idx : out_idx)[c].push_back((int)S.size() - 1);
        }
    }
    // (the vectors do not move any more: take the pointers now)
    std::vector<size_t> in_bytes(src_in.size()), out_bytes(src_out.size());
    for (int c = 0; c < nchunks; c++) {
        for (int i : in_idx[c]) { in_bytes[i] = src_in[i].mem.size() - 64; crew.in_jobs[c].push_back({dst_in[i].p(), src_in[i].p(), in_bytes[i]}); }
        for (int i : out_idx[c]) { out_bytes[i] = src_out[i].mem.size() - 64; crew.out_jobs[c].push_back({dst_out[i].p(), src_out[i].p(), out_bytes[i]}); }
    }
    crew.K = K; crew.nchunks = nchunks;
    pool.launch(&crew);
    for (int c = 0; c < nchunks; c++) {
        crew.inputs_of(c);
        for (int i : in_idx[c])
            if (memcmp(dst_in[i].p(), src_in[i].p(), in_bytes[i]) != 0) { printf("input of chunk %d not in place\n", c); return 1; }
        if (abort_early && c == nchunks / 2) return 0;   // ~BounceCrew: abort + wait for the workers to leave
    }
    for (int c = 0; c < nchunks; c++) {
        if (rng() % 4 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rnd(1, 300)));
        crew.landed(c);
    }
    crew.finish();
    for (size_t i = 0; i < src_out.size(); i++)
        if (memcmp(dst_out[i].p(), src_out[i].p(), out_bytes[i]) != 0) { printf("output %zu wrong\n", i); return 1; }
    // nothing wrote outside its job: the guard bytes around every destination are intact
    for (auto* V : {&dst_in, &dst_out})
        for (auto& b : *V) {
            const size_t n = b.mem.size() - 64;
            for (size_t i = 0; i < b.off; i++) if (b.mem[i] != 0) { printf("write before a destination\n"); return 1; }
            for (size_t i = b.off + n; i < b.mem.size(); i++) if (b.mem[i] != 0) { printf("write behind a destination\n"); return 1; }
        }
    return 0;
}

int main(int argc, char** argv) {
    const bool small = argc > 1;                            // (under the thread sanitizer: the same cases on less data)
    {
        CopyPool pool;
        const int Ks[] = {1, 3, 8, 16}, Cs[] = {1, 5, 32};
        for (int K : Ks) for (int C : Cs) for (int rep = 0; rep < (small ? 1 : 3); rep++)
            if (run_case(pool, K, C, small ? 20000 : (rep == 0 ? 3u << 20 : 70000), false)) { printf("FAILED K=%d chunks=%d\n", K, C); return 1; }
        for (int i = 0; i < (small ? 60 : 300); i++)        // one pool, many crews, changing thread counts, early returns in between
            if (run_case(pool, (int)rnd(1, 6), (int)rnd(1, 9), 5000, i % 7 == 3)) { printf("FAILED reuse %d\n", i); return 1; }
    }                                                       // ~CopyPool joins its threads
    { CopyPool idle; }                                      // a pool that never ran
    printf("ok\n");
    return 0;
}
"""


def test_copy_pool_and_crew_on_the_cpu(tmp_path):
    src = tmp_path / "harness.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "harness"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-Wall", "-I" + os.path.join(ROOT, "ilqg-mujoco_b200", "csrc"), str(src), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.returncode, r.stdout[-500:], r.stderr[-500:])


def test_copy_pool_under_thread_sanitizer(tmp_path):
    """The same harness under -fsanitize=thread (data races between the crew's counters, the jobs and the pool's hand-over)."""
    src = tmp_path / "harness.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "harness_tsan"
    c = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-fsanitize=thread", "-I" + os.path.join(ROOT, "ilqg-mujoco_b200", "csrc"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    if c.returncode != 0:
        pytest.skip("no thread sanitizer runtime in this toolchain: " + c.stderr[-200:])
    r = subprocess.run([str(exe), "small"], capture_output=True, text=True, timeout=600, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
    if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("thread sanitizer cannot map its shadow memory in this container")
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.returncode, r.stdout[-500:], r.stderr[-1500:])
