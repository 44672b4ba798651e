#!/bin/bash
# static SASS instruction count of one kernel of an object file: tools/sass_count.sh <object> <kernel-name-substring> [dump-file]
obj=$1; k=$2
cuobjdump -sass "$obj" | awk -v k="$k" '/Function : /{p=index($0,k)>0} p' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*//' > "${3:-/tmp/sass_count.txt}"
wc -l < "${3:-/tmp/sass_count.txt}"
