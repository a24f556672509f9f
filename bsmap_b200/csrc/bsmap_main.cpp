// bsmap -- drop-in command line; all work happens in libbsmap_b200.so (bsx_cli_main).
#include <cstdio>
#include <unistd.h>
extern "C" int bsx_cli_main(int argc, char **argv);
extern "C" void bsx_cli_exit_after_main(int on);
int main(int argc, char **argv) {
    bsx_cli_exit_after_main(1);
    const int rc = bsx_cli_main(argc, argv);     // outputs are flushed and closed when it returns
    fflush(stdout); fflush(stderr);
    _exit(rc);                                   // no device / context teardown: the driver reclaims everything with the process
}
