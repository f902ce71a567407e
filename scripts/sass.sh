#!/bin/bash
# scripts/sass.sh <substring of mangled kernel name> : SASS of the first matching kernel in the library
LIB=stain2stain_b200/lib/libs2s_b200.so
cuobjdump -sass $LIB | awk -v pat="$1" '
/Function :/ { show = (index($0, pat) > 0) }
show { print }'
