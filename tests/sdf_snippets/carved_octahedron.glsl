// octahedron with a cylinder carved out; helper with inout / out parameters, asin / tan, vec4 swizzles
void fold(inout vec3 p, out float m) {
    p = abs(p);
    m = (p.x + p.y + p.z - 0.7) * 0.57735027;
}

float sdf(in vec3 p) {
    vec4 q4 = vec4(p, 1.0);
    vec3 q = q4.xyz;
    float m;
    fold(q, m);
    float cyl = length(q4.wzyx.zw) - (0.18 + 0.05 * tan(0.4) + 0.02 * asin(0.5));
    return max(m, -cyl);
}

float sdfmaterial(in vec3 p) {
    return 1.0;
}
