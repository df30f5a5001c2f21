// Host-link probe: what can N GPUs of this box move to / from pinned host memory AT THE SAME TIME?
// The ceiling of bench.py's e2e leg at N GPUs (a product ciphertext is 1.27 MB and has to reach the host).
//
//   hostlink_probe [--mode proc|thread] [--dir d2h|h2d|both] [--mb 1024] [--iters 12] [--wc] [--bind] --set 0 --set 0,1 --set 0,1,2,3 ...
//
// Every worker (one process per GPU in `proc` mode - forked before any CUDA call - or one thread per GPU in `thread` mode)
// owns one device buffer and one pinned host buffer of --mb MiB, warms up, meets the others at a barrier in shared
// memory, then runs --iters cudaMemcpyAsync copies back to back on its own stream. Aggregate GB/s = all bytes moved /
// (latest end - earliest start) on CLOCK_MONOTONIC. --wc allocates the host buffer write-combined, --bind pins the worker
// to the CPUs the kernel lists as local to its GPU (/sys/bus/pci/devices/<id>/local_cpulist) before it allocates.
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct Shared {
    std::atomic<int> arrived;
    std::atomic<int> failed;
    double t0[16], t1[16];
    double bytes[16];
};

static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

struct Opt {
    std::string mode = "proc", dir = "d2h";
    size_t mb = 1024;
    int iters = 12;
    double secs = 0;      // > 0: instead of --iters copies, keep copying (one copy in flight, sync after each) until this much time has passed:
                          //      every worker is busy over the same window, so the sum is the steady-state aggregate
    bool wc = false, bind = false;
    int relay = 0;        // 1: worker for GPU d drives the copy from a stream of GPU (d + n/2) % n for d < n/2 (that GPU's copy engine reads d's
                          //    memory over NVLink and writes the host over ITS OWN PCIe link); 2: the same through a staging buffer on the relay GPU
    int relay_base = -1;  // first relay GPU (default: upper half of the device list of the set)
    std::vector<std::vector<int>> sets;
};

static void bind_near(int dev) {
    char bus[64];
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) return;
    for (char* p = bus; *p; p++) *p = (char)tolower(*p);
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return;
    char line[4096] = {0};
    if (!fgets(line, sizeof line, f)) { fclose(f); return; }
    fclose(f);
    cpu_set_t set;
    CPU_ZERO(&set);
    int any = 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a, b;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; c++) { CPU_SET(c, &set); any = 1; } }
        else if (sscanf(tok, "%d", &a) == 1) { CPU_SET(a, &set); any = 1; }
    }
    if (any) sched_setaffinity(0, sizeof set, &set);
}

static void worker(const Opt& o, Shared* sh, int slot, int nworkers, int dev, int relay_dev) {
    auto fail = [&](const char* what, cudaError_t e) {
        fprintf(stderr, "worker %d (gpu %d): %s: %s\n", slot, dev, what, cudaGetErrorString(e));
        sh->failed.fetch_add(1);
        sh->arrived.fetch_add(1);
    };
    cudaError_t e = cudaSetDevice(dev);
    if (e != cudaSuccess) return fail("cudaSetDevice", e);
    if (o.bind) bind_near(dev);
    const size_t bytes = o.mb << 20;
    void *d = nullptr, *h = nullptr, *d2 = nullptr, *h2 = nullptr;
    const bool both = o.dir == "both";
    if ((e = cudaMalloc(&d, bytes)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaHostAlloc(&h, bytes, o.wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc", e);
    if (both) {
        if ((e = cudaMalloc(&d2, bytes)) != cudaSuccess) return fail("cudaMalloc", e);
        if ((e = cudaHostAlloc(&h2, bytes, cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc", e);
        memset(h2, 1, bytes);
    }
    cudaMemset(d, 1, bytes);
    cudaDeviceSynchronize();
    void* stage = nullptr;
    if (relay_dev >= 0) {                      // everything below is issued with the relay GPU current
        if ((e = cudaSetDevice(relay_dev)) != cudaSuccess) return fail("cudaSetDevice(relay)", e);
        e = cudaDeviceEnablePeerAccess(dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess", e);
        cudaGetLastError();
        if (o.relay == 2 && (e = cudaMalloc(&stage, bytes)) != cudaSuccess) return fail("cudaMalloc(stage)", e);
    }
    cudaStream_t s, s2;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    const bool h2d = o.dir == "h2d";
    auto go = [&]() {
        if (h2d) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);
        else if (stage) {
            cudaMemcpyPeerAsync(stage, relay_dev, d, dev, bytes, s);
            cudaMemcpyAsync(h, stage, bytes, cudaMemcpyDeviceToHost, s);
        } else cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
        if (both) cudaMemcpyAsync(d2, h2, bytes, cudaMemcpyHostToDevice, s2);
    };
    for (int i = 0; i < 2; i++) go();
    cudaStreamSynchronize(s);
    cudaStreamSynchronize(s2);
    sh->arrived.fetch_add(1);
    while (sh->arrived.load() < nworkers) { /* spin: all workers start together */ }
    const double t0 = now();
    int done = 0;
    if (o.secs > 0) {
        while (now() - t0 < o.secs) { go(); cudaStreamSynchronize(s); done++; }
    } else {
        for (int i = 0; i < o.iters; i++) go();
        done = o.iters;
    }
    e = cudaStreamSynchronize(s);
    cudaError_t e2 = cudaStreamSynchronize(s2);
    const double t1 = now();
    if (e != cudaSuccess || e2 != cudaSuccess) { sh->failed.fetch_add(1); }
    sh->t0[slot] = t0;
    sh->t1[slot] = t1;
    sh->bytes[slot] = (double)bytes * done * (both ? 2 : 1);
    cudaFree(d); cudaFreeHost(h);
    if (both) { cudaFree(d2); cudaFreeHost(h2); }
}

static void run_set(const Opt& o, const std::vector<int>& devs) {
    Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    new (sh) Shared();
    sh->arrived = 0; sh->failed = 0;
    const int n = (int)devs.size();
    // relay assignment: with --relay-base B worker i uses GPU B + i % (number of relay GPUs = 8 - B); otherwise the lower half of the set uses the upper half
    auto relay_of = [&](int i) -> int {
        if (!o.relay) return -1;
        if (o.relay_base >= 0) return o.relay_base + i % (8 - o.relay_base) == devs[i] ? -1 : o.relay_base + i % (8 - o.relay_base);
        return i < n / 2 ? devs[i + n / 2] : -1;
    };
    if (o.mode == "proc") {
        std::vector<pid_t> kids;
        for (int i = 0; i < n; i++) {
            pid_t p = fork();
            if (p == 0) { worker(o, sh, i, n, devs[i], relay_of(i)); _exit(0); }
            kids.push_back(p);
        }
        for (pid_t p : kids) { int st; waitpid(p, &st, 0); }
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < n; i++) th.emplace_back(worker, std::cref(o), sh, i, n, devs[i], relay_of(i));
        for (auto& t : th) t.join();
    }
    double a = 1e300, b = 0, tot = 0;
    std::string per;
    for (int i = 0; i < n; i++) {
        a = sh->t0[i] < a ? sh->t0[i] : a;
        b = sh->t1[i] > b ? sh->t1[i] : b;
        tot += sh->bytes[i];
        char buf[64];
        snprintf(buf, sizeof buf, " %.1f", sh->bytes[i] / (sh->t1[i] - sh->t0[i]) / 1e9);
        per += buf;
    }
    std::string ds;
    for (int d : devs) ds += (ds.empty() ? "" : ",") + std::to_string(d);
    printf("{\"relay\": %d, \"relay_base\": %d, \"mode\": \"%s\", \"dir\": \"%s\", \"wc\": %d, \"bind\": %d, \"gpus\": [%s], \"n\": %d, \"mb\": %zu, \"iters\": %d, \"secs\": %.1f, \"aggregate_gbs\": %.1f, \"per_gpu_gbs\": [%s ], \"failed\": %d}\n",
           o.relay, o.relay_base, o.mode.c_str(), o.dir.c_str(), (int)o.wc, (int)o.bind, ds.c_str(), n, o.mb, o.iters, o.secs, tot / (b - a) / 1e9, per.c_str(), sh->failed.load());
    fflush(stdout);
    munmap(sh, sizeof(Shared));
}

int main(int argc, char** argv) {
    Opt o;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--mode") o.mode = val();
        else if (a == "--dir") o.dir = val();
        else if (a == "--mb") o.mb = (size_t)atol(val());
        else if (a == "--iters") o.iters = atoi(val());
        else if (a == "--secs") o.secs = atof(val());
        else if (a == "--wc") o.wc = true;
        else if (a == "--bind") o.bind = true;
        else if (a == "--relay") o.relay = atoi(val());
        else if (a == "--relay-base") o.relay_base = atoi(val());
        else if (a == "--set") {
            std::vector<int> s;
            std::string v = val();
            for (char* tok = strtok(&v[0], ","); tok; tok = strtok(nullptr, ",")) s.push_back(atoi(tok));
            o.sets.push_back(s);
        }
    }
    if (o.sets.empty()) o.sets.push_back({0});
    for (auto& s : o.sets) run_set(o, s);      // in `proc` mode the parent never touches CUDA, so every fork is clean
    return 0;
}
