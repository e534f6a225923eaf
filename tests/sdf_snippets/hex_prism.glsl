// hexagonal prism (after iquilezles): swizzle stores with compound assignment, sign(), rgba spellings
float sdf(in vec3 p) {
    const vec3 k = vec3(-0.8660254, 0.5, 0.57735);
    vec2 h = vec2(0.4, 0.25);
    p = abs(p);
    p.xy -= 2.0 * min(dot(k.xy, p.xy), 0.0) * k.xy;
    vec2 d = vec2(length(p.xy - vec2(clamp(p.x, -k.z * h.x, k.z * h.x), h.x)) * sign(p.y - h.x), p.z - h.y);
    return min(max(d.r, d.g), 0.0) + length(max(d, 0.0));
}

float sdfmaterial(in vec3 p) {
    return 1.0;
}
