#!/bin/bash
export SPARSE=1
N=130 ./gpu_var.sh dynenv_b200/libdynenv_b200.so tools/tmp/libinplace.so
