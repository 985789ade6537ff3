#!/usr/bin/env bash
# All-vs-all comparison of the metagenome samples in a directory -- same command line,
# outputs and resume behaviour as the reference workflow
# (bin/all_vs_all_metagenomes_IMSAME.sh:1-58):
#   for every unordered pair X < Y (directory listing order):
#     OUT/X-Y.align    IMSAME -query X -db Y
#     OUT/X-Y.r.align  IMSAME -query X -db revComp(Y)
#   existing outputs are kept (the run can be resumed); Y.r.EXT is a temporary.
# usage: all_vs_all_metagenomes_IMSAME.sh DIR COVERAGE SIMILARITY THREADS EXT OUTDIR
# IMSAME_EXTRA (optional env) is appended to every IMSAME call, e.g. "-gpus 8".
if [ $# != 6 ]; then
	echo "***ERROR*** Use: $0 metagenomes_directory coverage similarity threads file_extension outpath"
	exit -1
fi
dir=$1; cov=$2; sim=$3; thr=$4; ext=$5; out=$6
here="$( cd "$( dirname "${BASH_SOURCE[0]}" )" && pwd )"

samples=()
for f in $(ls -d "$dir"/*."$ext" | awk -F "/" '{print $NF}' | awk -F ".$ext" '{print $1}'); do
	samples+=("$f")
done

n=${#samples[@]}
for ((i = 0; i < n; i++)); do
	for ((j = i; j < n; j++)); do
		x=${samples[$i]}; y=${samples[$j]}
		if [ $i != $j ]; then
			if [[ ! -f $out/${x}-${y}.align ]]; then
				"$here"/IMSAME -query "$dir/$x.$ext" -db "$dir/$y.$ext" -n_threads "$thr" -coverage "$cov" -identity "$sim" -out "$out/${x}-${y}.align" $IMSAME_EXTRA
			fi
			if [[ ! -f $out/${x}-${y}.r.align ]]; then
				"$here"/revComp "$dir/$y.$ext" "$dir/$y.r.$ext"
				"$here"/IMSAME -query "$dir/$x.$ext" -db "$dir/$y.r.$ext" -n_threads "$thr" -coverage "$cov" -identity "$sim" -out "$out/${x}-${y}.r.align" $IMSAME_EXTRA
			fi
		fi
		# the reference removes the temporary on every iteration, including i == j (harmless error)
		rm "$dir/$y.r.$ext"
	done
done
