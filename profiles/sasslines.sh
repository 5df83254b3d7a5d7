#!/bin/bash
# Static SASS instruction count per source line of one kernel:  profiles/sasslines.sh OBJ.o kernel_name_substring [top]
set -e
obj=$1; kern=$2; top=${3:-40}
tmp=$(mktemp -d); (cd $tmp && cuobjdump -xelf all "$(realpath $obj)" >/dev/null)
nvdisasm -g -c $tmp/*.cubin | awk -v kern="$kern" '
/^\/\/-+ \.text\./ { infn = index($0, kern) > 0 }
/\/\/## File ".*", line [0-9]+/ { match($0,/line [0-9]+/); ln=substr($0,RSTART+5,RLENGTH-5); match($0,/File "[^"]*"/); f=substr($0,RSTART+6,RLENGTH-7); n=split(f,pp,"/"); key=pp[n]":"ln; next }
/^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@!]/ { if (infn) { cnt[key]++; tot++ } }
END { for (k in cnt) print cnt[k], k; print tot, "TOTAL" }' | sort -rn | head -$top
rm -rf $tmp
