// methratio_main.cpp -- the `methratio` executable: methratio.py's command line over libbsmap_b200.so
#include <cstdio>
#include <unistd.h>
extern "C" int bsx_methratio_main(int argc, char **argv);
int main(int argc, char **argv) {
    const int rc = bsx_methratio_main(argc, argv);   // the table is written and closed when it returns
    fflush(stdout); fflush(stderr);
    _exit(rc);
}
