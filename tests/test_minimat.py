"""Known answers for oracle/mlab/minimat.py, the mini-MATLAB interpreter that executes the reference's .m files (test
infrastructure).  Expected values are MATLAB's documented semantics (column-major linear indexing, 1-based subscripts, `end`,
colon, logical masks, implicit expansion, operator precedence, nested-function workspaces, multiple return values, meshgrid /
repmat / cat / circshift / sum / mean / min / max / sort / repelem orientation rules), each small enough to check by hand."""
import os

import numpy as np
import pytest

from oracle.mlab.minimat import Interp, MatlabError

PROGRAMS = {
    "idx": """
function [a,b,c,d,e,f,g,h] = idx()
A = [1 2 3; 4 5 6];            % 2 x 3
a = A(2);                      % column-major linear index -> 4
b = A(end);                    % 6
c = A(end,1);                  % 4
d = A(:,end-1);                % [2;5]
e = A(:);                      % [1;4;2;5;3;6]
B = zeros(2,3,2); B(:,:,2) = A; B(1,2,1) = 7;
f = B(:,:,2);
g = B(2,5);                    % trailing dimensions fold: column 5 = (2,:,2) -> B(2,2,2) = 5
M = A > 2 & A ~= 5;
A(M) = 0;
h = A;                         % [1 2 0; 0 5 0]
end
""",
    "prec": """
function [a,b,c,d,e,f,g] = prec()
a = -2^2;                      % -4
b = 2^-1;                      % 0.5
c = [1 -2];                    % two elements
d = [1 - 2];                   % one element
e = 1:3';                      % transpose binds tighter than colon -> 1 2 3
f = [1 2 3]';                  % column
x = [1 2 3];
g = [x', x'];                  % 3 x 2, the quote after a space inside brackets following an operand-less position is a transpose
end
""",
    "expand": """
function [a,b,c,d] = expand()
r = [1 2 3]; col = [10;20];
a = r + col;                   % 2 x 3 implicit expansion
w = zeros(1,1,2); w(1) = 1; w(2) = 3;
b = exp(w)./sum(exp(w));       % 1 x 1 x 2, sum along the first non-singleton dimension
c = repmat(w, 2, 2, 1);        % 2 x 2 x 2
d = size(c);
end
""",
    "shift": """
function [a,b,c,d,e,f] = shift()
A = [1 2 3; 4 5 6; 7 8 9];
a = circshift(A,-1);           % rows up: [4 5 6;7 8 9;1 2 3]
b = circshift(A,-1,2);         % columns left: [2 3 1;5 6 4;8 9 7]
c = circshift(A,1);            % rows down
d = cat(3, A, 2*A);
e = sum(sum(d(2:3,2:3,:),1),2);% 1 x 1 x 2: 28, 56
[X,Y] = meshgrid([1 2 3],[10 20]);
f = X + Y;                     % X(r,c) = x(c), Y(r,c) = y(r): [11 12 13; 21 22 23]
end
""",
    "nest": """
function [r, cnt, loc] = nest(n)
cnt = 0; acc = 0; k = 100;
for i = 1:n
    acc = acc + bump(i);
end
r = acc; loc = k;
    function y = bump(x)
        cnt = cnt + 1;         % shared with the parent (the parent uses cnt)
        t = x * 2;             % local to bump
        y = helper(t) + scale();
    end
    function y = helper(x)
        y = x + 1;
    end
    function s = scale()
        s = k / 100;           % reads the parent's k
    end
end
""",
    "multi": """
function [mn, mx, s, idx, n1, n2, z] = multi()
v = [3 1 2];
mn = min(v); mx = max(max(v), 2.5);
[s, idx] = sort(v, 'descend');
[n1, n2] = size(zeros(4, 5, 6));   % 4, 30
z = mean(mean([1 2; 3 4]));
end
""",
    "afun": """
function [p, q] = afun()
A = [1 2; 3 4]; off = 10;
[p, q] = arrayfun(@f2, A, [1 1; 2 2]);
    function [u, v] = f2(x, y)
        u = x * y + off;
        if x > 2, v = 1; elseif x > 1, v = 2; else, v = 3; end
    end
end
""",
    "afun3": """
function out = afun3()
S = cat(3, [1 2; 3 4], [10 20; 30 40]);   % 2 x 2 x 2 state next to a 2 x 2 index grid, as gqmap_gpu_mixture.m:29 does
out = arrayfun(@(s, g) s * 100 + g, S, [1 2; 3 4]);
end
""",
    "misc": """
function [a,b,c,d,e,f] = misc(o)
a = [o.dir,'/',num2str(12),'.png'];
b = repelem([1 2; 3 4], 2, 2);
c = mod(-1, 3);                 % 2
d = 0;
while 1
    d = d + 1;
    if d >= 5 || d < 0, break; end
end
e = min(max(7, 1), 5);
f = NaN(2,1); f(2) = 4;
end
""",
    "more": """
function [a,b,c,d,e,f,g,h] = more(varargin)
a = numel(varargin); b = varargin{2};
cellv = {'x', 'yz', 3};
c = [cellv{2}, cellv{1}];                 % 'yzx'
k = 10; add = @(x, y) x + y + k;          % captures k BY VALUE
k = 1000;
d = add(1, 2);                            % 13, not 1003
s = struct('p', 2, 'q', [1 2 3]);
s.r = s.p * 2;
e = s.r + s.q(end);                       % 7
G(:,:,2) = [1 2; 3 4];                    % created by indexed assignment, grows a zero first page
f = size(G); g = G(:,:,1);
name = 'dir/file.flo';
idx = strfind(name, '.'); idx = idx(end);
h = [name(idx:end), num2str(strcmp(name(idx:end), '.flo')), num2str(isfield(s, 'q')), num2str(ceil(2.1))];
end
""",
    "cleanup": """
function order = cleanup(n)
order = inner(n);
    function log = inner(n)
        log = 0;
        c1 = onCleanup(@() note(1));
        c2 = onCleanup(@() note(2));
        if n > 1, error('demo:fail', 'n was %d', n); end
        log = 10;
    end
    function note(k)
        order_sink(k);
    end
end
""",
}


@pytest.fixture(scope="module")
def interp(tmp_path_factory):
    root = tmp_path_factory.mktemp("mfiles")
    for name, src in PROGRAMS.items():
        if name == "cleanup":
            continue
        with open(os.path.join(root, name + ".m"), "w") as f:
            f.write(src.lstrip("\n"))
    return Interp([str(root)])


def eq(a, b):
    return np.array_equal(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))


def test_indexing(interp):
    a, b, c, d, e, f, g, h = interp.call("idx", nargout=8)
    assert (a, b, c, g) == (4.0, 6.0, 4.0, 5.0)
    assert eq(d, [[2], [5]]) and eq(e, [[1], [4], [2], [5], [3], [6]]) and eq(f, [[1, 2, 3], [4, 5, 6]])
    assert eq(h, [[1, 2, 0], [0, 5, 0]])


def test_precedence_and_brackets(interp):
    a, b, c, d, e, f, g = interp.call("prec", nargout=7)
    assert (a, b, d) == (-4.0, 0.5, -1.0)
    assert eq(c, [[1, -2]]) and eq(e, [[1, 2, 3]]) and eq(f, [[1], [2], [3]]) and eq(g, [[1, 1], [2, 2], [3, 3]])


def test_implicit_expansion_and_nd(interp):
    a, b, c, d = interp.call("expand", nargout=4)
    assert eq(a, [[11, 12, 13], [21, 22, 23]])
    assert b.shape == (1, 1, 2) and abs(b.sum() - 1) < 1e-15 and abs(b[0, 0, 1] / b[0, 0, 0] - np.exp(2)) < 1e-12
    assert c.shape == (2, 2, 2) and eq(c[:, :, 1], [[3, 3], [3, 3]]) and eq(d, [[2, 2, 2]])


def test_shifts_cat_meshgrid(interp):
    a, b, c, d, e, f = interp.call("shift", nargout=6)
    assert eq(a, [[4, 5, 6], [7, 8, 9], [1, 2, 3]]) and eq(b, [[2, 3, 1], [5, 6, 4], [8, 9, 7]]) and eq(c, [[7, 8, 9], [1, 2, 3], [4, 5, 6]])
    assert d.shape == (3, 3, 2) and eq(np.ravel(e), [28, 56]) and e.shape == (1, 1, 2)
    assert eq(f, [[11, 12, 13], [21, 22, 23]])


def test_nested_functions_share_the_parent_workspace(interp):
    r, cnt, loc = interp.call("nest", 3.0, nargout=3)
    assert cnt == 3.0 and loc == 100.0 and r == (2 + 1 + 1) + (4 + 1 + 1) + (6 + 1 + 1)


def test_multiple_outputs(interp):
    mn, mx, s, idx, n1, n2, z = interp.call("multi", nargout=7)
    assert (mn, mx, n1, n2, z) == (1.0, 3.0, 4.0, 30.0, 2.5) and eq(s, [[3, 2, 1]]) and eq(idx, [[1, 3, 2]])


def test_arrayfun_with_nested_handle(interp):
    p, q = interp.call("afun", nargout=2)
    assert eq(p, [[11, 12], [16, 18]]) and eq(q, [[3, 2], [1, 1]])


def test_arrayfun_expands_singleton_dimensions(interp):
    out = interp.call("afun3")
    assert out.shape == (2, 2, 2) and eq(out[:, :, 0], [[101, 202], [303, 404]]) and eq(out[:, :, 1], [[1001, 2002], [3003, 4004]])


def test_strings_loops_builtins(interp):
    a, b, c, d, e, f = interp.call("misc", {"dir": "out"}, nargout=6)
    assert a == "out/12.png" and c == 2.0 and d == 5.0 and e == 5.0
    assert eq(b, [[1, 1, 2, 2], [1, 1, 2, 2], [3, 3, 4, 4], [3, 3, 4, 4]]) and np.isnan(f[0, 0]) and f[1, 0] == 4.0


def test_cells_handles_structs_growth_strings(interp):
    a, b, c, d, e, f, g, h = interp.call("more", 1.0, 7.0, 3.0, nargout=8)
    assert (a, b, c, d, e) == (3.0, 7.0, "yzx", 13.0, 7.0)
    assert eq(f, [[2, 2, 2]]) and eq(g, [[0, 0], [0, 0]]) and h == ".flo113"


def test_oncleanup_runs_in_reverse_order_also_on_errors(tmp_path):
    with open(tmp_path / "cleanup.m", "w") as f:
        f.write(PROGRAMS["cleanup"].lstrip("\n"))
    seen = []
    I = Interp([str(tmp_path)], externals={"order_sink": lambda nargout, k: (seen.append(k), ())[1]})
    assert I.call("cleanup", 1.0) == 10.0 and seen == [2.0, 1.0]
    seen.clear()
    with pytest.raises(MatlabError) as e:
        I.call("cleanup", 5.0)
    assert e.value.ident == "demo:fail" and "n was 5" in str(e.value) and seen == [2.0, 1.0]


def test_scripts_and_file_io(tmp_path):
    with open(tmp_path / "job.m", "w") as f:
        f.write("clear;\nnames = {'a','b'};\nfor ti=1:numel(names)\n  opt.tag = names{ti}; opt.n = ti*2; % comment with 'quote\nend\n"
                "fid = fopen(out, 'w'); fwrite(fid, 'PIEH'); fwrite(fid, [7 9], 'int32'); fwrite(fid, [1.5; -2], 'float32'); fclose(fid);\n"
                "fid = fopen(out, 'r'); tag = fread(fid, 1, 'float32'); wh = fread(fid, 2, 'int32'); rest = fread(fid, inf, 'float32'); fclose(fid);\n"
                "save('x.mat', 'opt', 'wh');\n")
    I = Interp([str(tmp_path)])
    saved = []
    I.on_save = lambda fn, names, ws: saved.append((fn, names))
    target = str(tmp_path / "t.bin")
    # scripts share the caller's workspace in MATLAB; here the variable `out` is pre-seeded through a tiny wrapper script line
    with open(tmp_path / "job.m") as f:
        body = f.read()
    with open(tmp_path / "job.m", "w") as f:
        f.write("out = '%s';\n" % target + body.replace("clear;\n", ""))
    ws = I.run_script(str(tmp_path / "job.m"))
    assert ws["opt"] == {"tag": "b", "n": 4.0} and abs(ws["tag"] - 202021.25) < 1e-9 and eq(ws["wh"], [[7], [9]]) and eq(ws["rest"], [[1.5], [-2]])
    assert saved == [("x.mat", ("opt", "wh"))] and open(target, "rb").read()[:4] == b"PIEH"


def test_errors_are_loud(interp, tmp_path):
    with open(tmp_path / "bad.m", "w") as f:
        f.write("function y = bad()\nA = [1 2 3];\ny = A(4);\nend\n")
    with open(tmp_path / "bad2.m", "w") as f:
        f.write("function y = bad2()\ny = no_such_function(3);\nend\n")
    I = Interp([str(tmp_path)])
    with pytest.raises(MatlabError):
        I.call("bad")
    with pytest.raises(MatlabError):
        I.call("bad2")
