#!/bin/bash
# usage: tools/sass_of.sh <lib.so> <kernel-name-substring, e.g. render_kernelILi0ELi1ELb0ELb0E>  -> plain SASS listing on stdout
f=$(cuobjdump -elf "$1" 2>/dev/null | grep -o "_ZN[0-9A-Za-z_]*$2[0-9A-Za-z_]*" | head -1)
cuobjdump -sass -fun "$f" "$1" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
