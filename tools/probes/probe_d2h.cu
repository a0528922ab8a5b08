// probe_d2h.cu -- what can this box move from N GPUs into pinned host memory at once?  (development aid)
//
// The e2e leg of bench.py copies ~8 GB of observations per step and GPU to pinned host memory.  At N = 1 that runs at the
// PCIe ceiling (~55 GB/s); at N = 8 round 1 measured 92 GB/s AGGREGATE (11.5 GB/s per GPU).  This probe takes the engine
// out of the picture: one process per GPU (forked before any CUDA call, like torchrun's ranks), each copying `gib` GiB
// device -> host with bare cudaMemcpyAsync in 512 MiB pieces over two streams, all processes released together; the
// aggregate is total bytes / slowest process.  Variants of the host allocation:
//   hostalloc   cudaHostAlloc(default)                       -- what torch's pin_memory uses
//   wc          cudaHostAlloc(write-combined)
//   thp         mmap + madvise(MADV_HUGEPAGE) + first touch + cudaHostRegister  (2 MiB pages: fewer IOMMU entries)
//   numa        like thp, first-touched after binding the process to the CPUs of the GPU's NUMA node (sysfs), when the
//               box exposes more than one node
// usage: probe_d2h [gib_per_gpu=4] [max_gpus=8] [quick=0]   (quick: fewer combinations -- every GPU count costs box time)
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
#include <vector>

struct Shared {
    std::atomic<int> ready, go;
    double seconds[16];
    int error[16];
};

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int gpu_numa_node(int dev)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) return -1;
    for (char *p = bus; *p; ++p) *p = char(tolower(*p));
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

static bool bind_to_node_cpus(int node)
{
    if (node < 0) return false;
    char path[128];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    char buf[512] = {0};
    const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (char *tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int lo, hi;
        if (sscanf(tok, "%d-%d", &lo, &hi) == 2) { for (int c = lo; c <= hi; ++c) CPU_SET(c, &set); }
        else if (sscanf(tok, "%d", &lo) == 1) CPU_SET(lo, &set);
    }
    return sched_setaffinity(0, sizeof(set), &set) == 0;
}

enum Mode { HOSTALLOC, WC, THP, NUMA, N_MODES };
static const char *mode_name[N_MODES] = {"hostalloc", "wc", "thp", "numa"};

static int child(int rank, int n, int mode, size_t bytes, int direction, Shared *sh)
{
    if (cudaSetDevice(rank) != cudaSuccess) return 1;
    const size_t piece = 512ull << 20, buf = std::min(bytes, size_t(2) << 30);
    uint8_t *dev = nullptr, *host = nullptr;
    if (cudaMalloc(&dev, buf) != cudaSuccess) return 2;
    cudaMemset(dev, 1, buf);
    if (mode == HOSTALLOC || mode == WC) {
        if (cudaHostAlloc(&host, buf, mode == WC ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) return 3;
    } else {
        if (mode == NUMA) bind_to_node_cpus(gpu_numa_node(rank));
        void *p = mmap(nullptr, buf, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) return 4;
        madvise(p, buf, MADV_HUGEPAGE);
        host = static_cast<uint8_t *>(p);
        for (size_t i = 0; i < buf; i += 4096) host[i] = 0;  // first touch on this CPU's node
        if (cudaHostRegister(host, buf, cudaHostRegisterDefault) != cudaSuccess) return 5;
    }
    cudaStream_t s[2];
    cudaStreamCreateWithFlags(&s[0], cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s[1], cudaStreamNonBlocking);
    auto run = [&](size_t total) {
        size_t done = 0;
        int k = 0;
        while (done < total) {
            const size_t off = done % buf, nb = std::min(piece, std::min(total - done, buf - off));
            if (direction == 0) cudaMemcpyAsync(host + off, dev + off, nb, cudaMemcpyDeviceToHost, s[k & 1]);
            else cudaMemcpyAsync(dev + off, host + off, nb, cudaMemcpyHostToDevice, s[k & 1]);
            done += nb;
            ++k;
        }
        cudaStreamSynchronize(s[0]);
        cudaStreamSynchronize(s[1]);
    };
    run(buf);  // warm-up
    sh->ready.fetch_add(1);
    while (sh->go.load() == 0) usleep(200);
    const double t0 = now();
    run(bytes);
    sh->seconds[rank] = now() - t0;
    sh->error[rank] = cudaGetLastError() == cudaSuccess ? 0 : 9;
    return 0;
}

int main(int argc, char **argv)
{
    const double gib = argc > 1 ? atof(argv[1]) : 4.0;
    const int max_gpus = argc > 2 ? atoi(argv[2]) : 8;
    const bool quick = argc > 3 && atoi(argv[3]) != 0;
    const size_t bytes = size_t(gib * 1024.0 * 1024.0 * 1024.0);
    // the device count must be learnt WITHOUT creating a CUDA context in the parent (children fork below)
    int n_dev = 0;
    {
        int fd[2];
        if (pipe(fd) != 0) return 1;
        const pid_t pid = fork();
        if (pid == 0) {
            int n = 0;
            cudaGetDeviceCount(&n);
            if (write(fd[1], &n, sizeof(n)) != sizeof(n)) _exit(1);
            _exit(0);
        }
        if (read(fd[0], &n_dev, sizeof(n_dev)) != sizeof(n_dev)) n_dev = 0;
        waitpid(pid, nullptr, 0);
    }
    printf("devices: %d, %.1f GiB per GPU and run\n", n_dev, gib);
    Shared *sh = static_cast<Shared *>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
    for (int direction = 0; direction < 2; ++direction) {
        for (int n = 1; n <= std::min(n_dev, max_gpus); n *= 2) {
            for (int mode = 0; mode < N_MODES; ++mode) {
                if (direction == 1 && mode != HOSTALLOC && mode != THP) continue;
                const bool top = n * 2 > std::min(n_dev, max_gpus);  // the largest GPU count of the sweep
                if (quick && direction == 1 && !(top && mode == HOSTALLOC)) continue;
                if (quick && direction == 0 && (mode == WC || mode == NUMA) && !top) continue;
                new (sh) Shared();
                sh->ready = 0;
                sh->go = 0;
                std::vector<pid_t> kids;
                for (int r = 0; r < n; ++r) {
                    const pid_t pid = fork();
                    if (pid == 0) _exit(child(r, n, mode, bytes, direction, sh));
                    kids.push_back(pid);
                }
                const double t_wait = now();
                while (sh->ready.load() < n && now() - t_wait < 120.0) usleep(1000);
                sh->go = 1;
                int bad = 0;
                for (pid_t pid : kids) {
                    int status = 0;
                    waitpid(pid, &status, 0);
                    if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) bad = WIFEXITED(status) ? WEXITSTATUS(status) : -1;
                }
                double slowest = 0, fastest = 1e30;
                for (int r = 0; r < n; ++r) { slowest = std::max(slowest, sh->seconds[r]); fastest = std::min(fastest, sh->seconds[r]); }
                if (bad) printf("%s n=%d %-9s FAILED (child exit %d)\n", direction ? "H2D" : "D2H", n, mode_name[mode], bad);
                else
                    printf("%s n=%d %-9s aggregate %7.1f GB/s   per GPU %6.1f GB/s (slowest)  %6.1f GB/s (fastest)\n",
                           direction ? "H2D" : "D2H", n, mode_name[mode], n * double(bytes) / slowest / 1e9, double(bytes) / slowest / 1e9,
                           double(bytes) / fastest / 1e9);
                fflush(stdout);
            }
        }
    }
    return 0;
}
