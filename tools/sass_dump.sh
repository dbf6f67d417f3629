#!/bin/bash
# usage: tools/sass_dump.sh <object> <function-substring> -> one instruction per line (address opcode operands)
obj=$1; pat=$2
fn=$(cuobjdump -sass "$obj" | grep "Function :" | grep "$pat" | head -1 | sed 's/.*Function : //')
cuobjdump -sass -fun "$fn" "$obj" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#^\s+/\*([0-9a-f]+)\*/\s+#\1 #; s#/\*.*##'
