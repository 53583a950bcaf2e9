// LD_PRELOAD helper (debugging aid): native backtrace on SIGSEGV.
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void handler(int sig) {
    void* frames[64];
    int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    _exit(139);
}
__attribute__((constructor)) static void init(void) { signal(SIGSEGV, handler); }
