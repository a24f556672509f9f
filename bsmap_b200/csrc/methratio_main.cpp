// methratio_main.cpp -- the `methratio` executable: methratio.py's command line over libbsmap_b200.so
extern "C" int bsx_methratio_main(int argc, char **argv);
int main(int argc, char **argv) { return bsx_methratio_main(argc, argv); }
