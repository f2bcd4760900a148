# usage: sass_fn.sh lib.so mangled_name > out.sass
cuobjdump -sass "$1" | awk -v fn="$2" '/Function :/{p=($3==fn)} p'
