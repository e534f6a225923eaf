"""Builds the SDF dispatchers of one scene for the CPU oracle -- TEST INFRASTRUCTURE.

Independent restatement of what the reference's InsertSDF does to shader.comp (host:2004-2054) plus the token-level
GLSL -> C++ rewrite needed to compile the snippets with g++ against include/pt_glsl.h:
  * first "sdf" substring -> SDF{i+1}, then first "sdfmaterial" -> SDF{i+1}MATERIAL      (host:2015-2017)
  * dispatcher lines, SDF 1 first, the material line before the distance line            (host:2046-2051)
  * float literals get an `f` suffix, `in` qualifiers vanish, `out/inout T x` -> `T& x`, multi-letter swizzles
    become swizzle calls (`.sw3<0,2,1>()` reads, `.lsw2<0,2>()` in front of an assignment operator), array
    constructors `T[n](..)` become braced initialisers; `#define` lines go through as they are.
The product has its own C++ front-end (pathtracer_b200/csrc/pt_sdf_front.cpp); tests compare both on random points.
"""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, '_build')
CXX = '/usr/bin/g++' if os.access('/usr/bin/g++', os.X_OK) else 'g++'
CXXFLAGS = ['-O2', '-std=c++17', '-fPIC', '-shared', '-mfma', '-mavx2', '-ffp-contract=off', '-fno-fast-math',
            '-I' + os.path.join(HERE, '..', 'include')]

_FLOAT = re.compile(r'(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])')


def translate_snippet(glsl, number):
    s = glsl.replace('\r\n', '\n').replace('\r', '\n')
    i = s.find('sdf')
    if i < 0:
        raise ValueError('snippet %d has no "sdf"' % number)
    s = s[:i] + 'SDF%d' % number + s[i + 3:]
    i = s.find('sdfmaterial')
    if i < 0:
        raise ValueError('snippet %d has no "sdfmaterial"' % number)
    s = s[:i] + 'SDF%dMATERIAL' % number + s[i + 11:]
    s = re.sub(r'//[^\n]*', '', s)
    s = re.sub(r'/\*.*?\*/', '', s, flags=re.S)
    s = _FLOAT.sub(lambda m: m.group(1) + 'f', s)
    s = re.sub(r'\b(?:const\s+)?in\s+(?=(?:float|int|uint|bool|vec[234]|mat[234])\b)', '', s)
    s = re.sub(r'\b(?:inout|out)\s+(float|int|uint|bool|vec[234]|mat[234])\s+', r'\1& ', s)
    s = _array_constructors(s)
    # (the product's front end numbers its sin( sites for the fast build's FMA-pipe option; on the CPU and in strict
    # builds every site is sin(), so the oracle's translation leaves them alone)

    def swz(m):
        letters = m.group(1)
        for alphabet in ('xyzw', 'rgba', 'stpq'):
            if all(c in alphabet for c in letters):
                idx = ','.join(str(alphabet.index(c)) for c in letters)
                break
        else:
            return m.group(0)
        kind = 'lsw' if m.group(2) else 'sw'
        return '.%s%d<%s>()%s' % (kind, len(letters), idx, m.group(2) or '')
    s = re.sub(r'\.([xyzwrgbastpq]{2,4})\b(?!\s*\()(\s*(?:[-+*/]=|=(?!=)))?', swz, s)
    s = re.sub(r'\.([rgbastpq])\b(?!\s*\()', lambda m: '.' + 'xyzwxyzw'['rgbastpq'.index(m.group(1))], s)
    return s


def _array_constructors(s):
    """`float[3](a, b, c)` / `vec2[](u, v)` -> `{a, b, c}` (matching parenthesis found by depth counting)"""
    out, i = [], 0
    rx = re.compile(r'\b(?:float|int|uint|bool|vec[234]|mat[234])\s*\[\s*\d*\s*\]\s*\(')
    while True:
        m = rx.search(s, i)
        if not m:
            out.append(s[i:])
            return ''.join(out)
        depth, j = 1, m.end()
        while depth and j < len(s):
            depth += {'(': 1, ')': -1}.get(s[j], 0)
            j += 1
        out.append(s[i:m.start()] + '{' + _array_constructors(s[m.end():j - 1]) + '}')
        i = j


PRELUDE = r'''
#include "pt_glsl.h"
namespace ptglsl {
static const float* sdfs;
/* shader.comp:681-683 */
inline float minMaterial(float x, float y, float material1, float material2) { return (x < y) ? material1 : material2; }
/* shader.comp:686-693 */
inline float smin(float x, float y) {
    float k = 0.02f;
    k *= 6.0f;
    float h = max(k - abs(x - y), 0.0f) / k;
    float m = h * h * h * 0.5f;
    float s = m * k * ONEBYTHREE;
    return min(x, y) - s;
}
/* shader.comp:695-702 */
inline vec2 smin(vec2 x, vec2 y) {
    float k = 0.02f;
    k *= 6.0f;
    float h = max(k - abs(x.x - y.x), 0.0f) / k;
    float m = h * h * h * 0.5f;
    float s = m * k * ONEBYTHREE;
    return (x.x < y.x) ? vec2(x.x - s, x.y + (y.y - x.y) * m) : vec2(y.x - s, x.y + (y.y - x.y) * (1.0f - m));
}
'''


def generate(sources):
    out = [PRELUDE]
    for i, src in enumerate(sources):
        out.append(translate_snippet(src, i + 1))
    sdf_lines, mat_lines = [], []
    for i in range(len(sources)):
        code = 1 << (i % 32)
        word = i // 32 + 1          # InsertSDF: "(set" + ((i - i % 32) / 32 + 1) + " & " + 2^(i % 32)  (host:2012, 2029-2033)
        if word > 4:
            raise ValueError('more than 128 SDFs')
        pos = '(p - vec3(sdfs[%d], sdfs[%d], sdfs[%d]))' % (6 * i, 6 * i + 1, 6 * i + 2)
        cond = 'if ((set%d & %du) == %du) ' % (word, code, code)
        sdf_line = cond + 'sdf = min(sdf, SDF%d%s);' % (i + 1, pos)
        mat_line = cond + 'sdfmaterial = minMaterial(sdf, SDF%d%s, sdfmaterial, SDF%dMATERIAL%s);' % (i + 1, pos, i + 1, pos)
        sdf_lines.append(sdf_line)
        mat_lines += [mat_line, sdf_line]
    out.append('''
/* shader.comp:706-711 */
inline float SDF(vec3 p, uint set1, uint set2, uint set3, uint set4) {
    float sdf = MAXDIST;
    %s
    return sdf;
}
/* shader.comp:713-719 */
inline float SDFMATERIAL(vec3 p, uint set1, uint set2, uint set3, uint set4) {
    float sdf = MAXDIST;
    float sdfmaterial = 0.0f;
    %s
    return sdfmaterial;
}
} // namespace ptglsl
extern "C" float oracle_SDF4(const float* s, float x, float y, float z, unsigned set1, unsigned set2, unsigned set3, unsigned set4) {
    ptglsl::sdfs = s;
    return ptglsl::SDF(ptglsl::vec3(x, y, z), set1, set2, set3, set4);
}
extern "C" float oracle_SDFMATERIAL4(const float* s, float x, float y, float z, unsigned set1, unsigned set2, unsigned set3, unsigned set4) {
    ptglsl::sdfs = s;
    return ptglsl::SDFMATERIAL(ptglsl::vec3(x, y, z), set1, set2, set3, set4);
}
extern "C" float oracle_SDF(const float* s, float x, float y, float z, unsigned set1) { return oracle_SDF4(s, x, y, z, set1, 0u, 0u, 0u); }
extern "C" float oracle_SDFMATERIAL(const float* s, float x, float y, float z, unsigned set1) { return oracle_SDFMATERIAL4(s, x, y, z, set1, 0u, 0u, 0u); }
''' % ('\n    '.join(sdf_lines), '\n    '.join(mat_lines)))
    return '\n'.join(out)


def build(sources):
    """Returns the path of a shared object exporting oracle_SDF / oracle_SDFMATERIAL ('' when no SDF)."""
    if not sources:
        return ''
    text = generate(sources)
    tag = hashlib.sha1(text.encode()).hexdigest()[:16]
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, 'sdf_%s.so' % tag)
    if not os.path.exists(so):
        cpp = os.path.join(BUILD, 'sdf_%s.cpp' % tag)
        with open(cpp, 'w') as f:
            f.write(text)
        tmp = so + '.tmp%d' % os.getpid()
        subprocess.run([CXX, *CXXFLAGS, '-o', tmp, cpp], check=True)
        os.replace(tmp, so)
    return so
