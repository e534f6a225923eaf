// domain repetition through a swizzle store: p.xz = mod(..)
float boxDist(in vec3 p, in vec3 b) {
    vec3 q = abs(p) - b;
    return length(max(q, 0.0)) + min(max(q.x, max(q.y, q.z)), 0.0);
}

float sdf(in vec3 p) {
    const vec2 c = vec2(0.5, 0.7);
    p.xz = mod(p.xz + 0.5 * c, c) - 0.5 * c;
    p.y -= 0.1;
    return boxDist(p, vec3(0.12, 0.3, 0.15)) - 0.02;
}

float sdfmaterial(in vec3 p) {
    return 0.0;
}
