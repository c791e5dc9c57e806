#!/bin/bash
# usage: tools/sass_of.sh <lib> <substring of the mangled kernel name>  -> SASS of the matching functions (no encodings)
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{on = index($0, pat) > 0} on' | grep -v "^\s*/\* 0x"
