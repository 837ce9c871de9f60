/* LD_PRELOAD diagnostic: a SIGSEGV/SIGBUS/SIGABRT handler on an alternate stack that prints the faulting address, the thread id,
 * a backtrace (return addresses + module offsets) and /proc/self/maps to stderr, then re-raises with the default action.
 * Build: gcc -O1 -g -shared -fPIC -o gpurun_out/segv_trace.so tools/probe/segv_trace.c
 * Use:   LD_PRELOAD=gpurun_out/segv_trace.so oracle/_ref/rayforce_ref -f integration/demo/plugin.rfl */
#define _GNU_SOURCE
#include <execinfo.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <ucontext.h>
#include <unistd.h>

static char alt[1 << 16];
static void on_fault(int sig, siginfo_t *si, void *uc_) {
    ucontext_t *uc = (ucontext_t *)uc_;
    char line[256];
    void *frames[64];
    int n = snprintf(line, sizeof line, "\n[segv_trace] signal %d addr %p tid %ld rip %p rsp %p\n", sig, si->si_addr, (long)syscall(SYS_gettid),
                     (void *)uc->uc_mcontext.gregs[REG_RIP], (void *)uc->uc_mcontext.gregs[REG_RSP]);
    write(2, line, n);
    n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    int fd = open("/proc/self/maps", O_RDONLY);
    if (fd >= 0) {
        char buf[4096];
        ssize_t r;
        write(2, "[segv_trace] maps:\n", 19);
        while ((r = read(fd, buf, sizeof buf)) > 0) write(2, buf, r);
        close(fd);
    }
    signal(sig, SIG_DFL);
    raise(sig);
}
__attribute__((constructor)) static void install(void) {
    stack_t ss = {.ss_sp = alt, .ss_size = sizeof alt, .ss_flags = 0};
    sigaltstack(&ss, NULL);
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = on_fault;
    sa.sa_flags = SA_SIGINFO | SA_ONSTACK | SA_RESETHAND;
    sigaction(SIGSEGV, &sa, NULL);
    sigaction(SIGBUS, &sa, NULL);
    sigaction(SIGABRT, &sa, NULL);
    void *f[1];
    backtrace(f, 1);   /* load libgcc now, not inside the handler */
}
