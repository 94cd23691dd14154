#!/bin/bash
export SPARSE=1
N=130 ./gpu_var.sh dynenv_b200/libdynenv_b200.so tools/tmp/libs_nobar1.so tools/tmp/libs_nobar2.so tools/tmp/libs_nobar12.so
