// bsmap -- drop-in command line; all work happens in libbsmap_b200.so (bsx_cli_main).
extern "C" int bsx_cli_main(int argc, char **argv);
int main(int argc, char **argv) { return bsx_cli_main(argc, argv); }
