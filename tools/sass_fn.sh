#!/bin/bash
# usage: tools/sass_fn.sh <mangled-name-substring>  -> SASS of the first matching function in libmgym.so
cuobjdump -sass modurl_gym_b200/libmgym.so 2>/dev/null | awk -v pat="$1" '/Function : /{f=index($0,pat)>0} f{print}' | grep -E "^\s+/\*[0-9a-f]{4,6}\*/" | sed -e 's/ *\/\* 0x[0-9a-f]* \*\///'
