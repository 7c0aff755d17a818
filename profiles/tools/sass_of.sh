#!/bin/bash
# usage: sass_of.sh <lib.so> <mangled-substring>  -> SASS of that kernel with source lines (stdout)
tmp=$(mktemp -d); cd $tmp
cuobjdump -xelf all "$1" > /dev/null
nvdisasm -g -c *.cubin | awk -v k="$2" '/^\.text\./{p=index($0,k)>0} /^\.section/{p=0} p'
rm -rf $tmp
