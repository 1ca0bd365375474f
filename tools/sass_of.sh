#!/bin/bash
# usage: tools/sass_of.sh <lib.so> <function-substring>   -> prints the SASS of the first matching function
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/ {on = index($0, pat) > 0} on {print}'
