#!/bin/bash
# usage: tools/sass_of.sh <object-or-so> <mangled-name-substring>  -> cleaned SASS of the first matching kernel
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function : /{on = index($0, pat) > 0} on {print}' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+//; s/\s+\/\* 0x[0-9a-f]+ \*\/$//'
