#!/bin/bash
export SPARSE=1
N=130 ./gpu_var.sh dynenv_b200/libdynenv_b200.so tools/tmp/libs_w16b1.so tools/tmp/libs_w8b3.so tools/tmp/libs_w8b2.so tools/tmp/libs_w8b1.so
