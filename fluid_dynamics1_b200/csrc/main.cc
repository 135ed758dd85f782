// cnavier_b200: `./cnavier_b200 [config_file] [output_folder]`, the reference's command line
// (src/main.c:37-89) on the device-resident path.
#include "../../include/cnavier_b200.h"
int main(int argc, char **argv) { return cnv_main(argc, argv); }
