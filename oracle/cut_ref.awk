# oracle/cut_ref.awk — cuts ONE function definition out of a reference source file, verbatim, at build time.
#   awk -v start='<literal prefix of the signature line, after leading blanks>' -f cut_ref.awk file.cc
# Prints from the first line whose stripped text begins with `start` through the first later line that is a lone `}` at the
# signature's own indentation.  The output goes to oracle/_ref/ (git-ignored): reference text is compiled, never committed.
{
    line = $0
    stripped = line; sub(/^[ \t]+/, "", stripped)
    if (!on && index(stripped, start) == 1) { on = 1; indent = substr(line, 1, length(line) - length(stripped)) }
    if (on) {
        print line
        if (line == indent "}") { done = 1; exit }
    }
}
END { if (!done) { print "cut_ref.awk: no match for: " start > "/dev/stderr"; exit 1 } }
