// measurement only: can the copy engine read page-cache pages directly (mmap + cudaHostRegister) faster than pread -> pinned -> DMA?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
using clk = std::chrono::steady_clock;
static double ms(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count() * 1e3; }
int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "/tmp/hostreg.bin";
    size_t n = (size_t)256 << 20;
    { int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644); char* b = (char*)malloc(1 << 20); memset(b, 'C', 1 << 20); for (size_t i = 0; i < n; i += 1 << 20) if (write(fd, b, 1 << 20) < 0) return 1; close(fd); free(b); }
    char* d; cudaMalloc(&d, n); cudaStream_t st; cudaStreamCreate(&st);
    char* pin; cudaHostAlloc(&pin, n, cudaHostAllocDefault);
    cudaMemcpyAsync(d, pin, n, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st);
    for (int rep = 0; rep < 3; rep++) {
        int fd = open(path, O_RDONLY);
        auto t0 = clk::now();
        char* m = (char*)mmap(nullptr, n, PROT_READ, MAP_SHARED | MAP_POPULATE, fd, 0);
        auto t1 = clk::now();
        cudaError_t e = cudaHostRegister(m, n, cudaHostRegisterReadOnly);
        auto t2 = clk::now();
        if (e != cudaSuccess) { printf("cudaHostRegister(ReadOnly): %s\n", cudaGetErrorString(e)); cudaGetLastError(); }
        else {
            cudaMemcpyAsync(d, m, n, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st);
            auto t3 = clk::now();
            cudaHostUnregister(m);
            auto t4 = clk::now();
            printf("mmap+populate %.2f  register %.2f  copy %.2f (%.1f GB/s)  unregister %.2f ms  | total %.2f ms for %zu MB\n", ms(t0, t1), ms(t1, t2), ms(t2, t3), n / ms(t2, t3) / 1e6, ms(t3, t4), ms(t0, t4), n >> 20);
        }
        // pageable copy straight from the mapping (driver bounce buffers)
        auto t5 = clk::now();
        cudaMemcpyAsync(d, m, n, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st);
        auto t6 = clk::now();
        printf("pageable copy from the mapping %.2f ms (%.1f GB/s)\n", ms(t5, t6), n / ms(t5, t6) / 1e6);
        munmap(m, n); close(fd);
        // single-thread pread -> pinned, then copy
        fd = open(path, O_RDONLY);
        auto t7 = clk::now();
        size_t a = 0; while (a < n) { ssize_t g = pread(fd, pin + a, n - a, a); if (g <= 0) break; a += g; }
        auto t8 = clk::now();
        cudaMemcpyAsync(d, pin, n, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st);
        auto t9 = clk::now();
        printf("1-thread pread %.2f ms (%.1f GB/s), pinned copy %.2f ms (%.1f GB/s)\n", ms(t7, t8), n / ms(t7, t8) / 1e6, ms(t8, t9), n / ms(t8, t9) / 1e6);
        close(fd);
    }
    unlink(path);
    return 0;
}
