#!/bin/sh
# usage: tools/sass_of.sh <substring of mangled kernel name>  -> address + instruction listing of that kernel
cuobjdump -sass "$(dirname "$0")/../snpgenie_b200/libsnpgenie_cuda.so" | awk -v pat="$1" '/Function :/{p=index($0,pat)>0;next} p' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
