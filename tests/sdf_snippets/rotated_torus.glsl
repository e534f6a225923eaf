// torus, rotated about y by a mat2 applied to a swizzle
float torus(in vec3 p, in vec2 t) {
    vec2 q = vec2(length(p.xz) - t.x, p.y);
    return length(q) - t.y;
}

float sdf(in vec3 p) {
    float a = 0.6;
    p.xz *= mat2(cos(a), -sin(a), sin(a), cos(a));
    p.xy *= mat2(cos(0.3), -sin(0.3), sin(0.3), cos(0.3));
    return torus(p, vec2(0.6, 0.2));
}

float sdfmaterial(in vec3 p) {
    return 1.0;
}
